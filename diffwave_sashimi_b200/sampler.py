"""Diffusion schedule and the reverse sampler with the reference's signatures
(utils.py:121-151 `calc_diffusion_hyperparams`, generate.py:23-55 `sampling`)."""
import torch


def calc_diffusion_hyperparams(T, beta_0, beta_T, beta=None, fast=False):
    """Same tables as the reference, computed with the same fp32 torch ops in the same order so
    they are bit-identical; `Sigma` stays on the CPU like the reference's (utils.py:150)."""
    if fast and beta is not None:
        Beta = torch.tensor(beta)
        T = len(beta)
    else:
        Beta = torch.linspace(beta_0, beta_T, T)
    Alpha = 1 - Beta
    Alpha_bar = Alpha + 0
    Beta_tilde = Beta + 0
    for t in range(1, T):
        Alpha_bar[t] *= Alpha_bar[t - 1]
        Beta_tilde[t] *= (1 - Alpha_bar[t - 1]) / (1 - Alpha_bar[t])
    Sigma = torch.sqrt(Beta_tilde)
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    return {"T": T, "Beta": Beta.to(dev), "Alpha": Alpha.to(dev), "Alpha_bar": Alpha_bar.to(dev), "Sigma": Sigma}


def step_coefficients(dh):
    """(3,T) host table for dwb_sample: (1-alpha)/sqrt(1-alpha_bar), sqrt(alpha), sigma — each
    formed by the fp32 torch expression the reference evaluates per step (generate.py:52-54)."""
    Alpha, Alpha_bar, Sigma = dh["Alpha"].cpu(), dh["Alpha_bar"].cpu(), dh["Sigma"].cpu()
    return torch.stack([(1 - Alpha) / torch.sqrt(1 - Alpha_bar), torch.sqrt(Alpha), Sigma]).float().contiguous()


def draw_noise(size, T, pin=True):
    """The reference's RNG consumption: x_T = torch.normal(0,1,size) first, then one draw per step
    for t = T-1 .. 1, all on the CPU default generator (generate.py:47,54; no draw at t = 0)."""
    x_T = torch.normal(0, 1, size=size)
    noise = torch.empty((max(T - 1, 0),) + tuple(size))
    for i in range(T - 1):
        noise[i] = torch.normal(0, 1, size=size)
    if pin and torch.cuda.is_available():
        x_T, noise = x_T.pin_memory(), noise.pin_memory()
    return x_T, noise


@torch.no_grad()
def sampling(net, size, diffusion_hyperparams, condition=None, use_graph=True, verbose=True):
    """Drop-in for generate.sampling(net, size, diffusion_hyperparams, condition): the T network
    evaluations and DDPM updates run as one CUDA graph inside libdwb."""
    dh = diffusion_hyperparams
    T = dh["T"]
    assert len(dh["Alpha"]) == T and len(dh["Alpha_bar"]) == T and len(dh["Sigma"]) == T and len(size) == 3
    if verbose:
        print("begin sampling, total number of reverse steps = %s" % T)
    eng = net._engine_get()
    x_T, noise = draw_noise(size, T)
    x_T = x_T.to(eng.device, non_blocking=True)
    noise = noise.to(eng.device, non_blocking=True)
    return eng.sample(x_T, noise, step_coefficients(dh), condition, use_graph=use_graph)
