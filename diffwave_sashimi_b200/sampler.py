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


class GlobalNoise:
    """The reference's noise source: the process-wide CPU generator, x_T first, then one (B,1,L) draw per step
    (generate.py:47,54).  `fill(out)` = `out <- torch.normal(0, 1, size=out.shape)` without the allocation:
    same generator, same kernel (contiguous float tensor of >= 16 elements), bit-identical values."""

    def fill(self, out):
        out.normal_()


class PerClipNoise:
    """One seeded CPU generator per clip: clip c of the batch sees exactly the stream `sampling(net, (1,1,L))`
    would consume after `torch.manual_seed(seeds[c])` - so a clip's audio does not depend on which other clips
    share its batch or on how a global batch is sharded over GPUs, and a rank draws only its own clips."""

    def __init__(self, seeds):
        self.gens = [torch.Generator().manual_seed(int(s)) for s in seeds]

    def fill(self, out):
        assert out.shape[0] == len(self.gens)
        for c, g in enumerate(self.gens):
            out[c].normal_(generator=g)


def clip_seed(seed, clip):
    """Seed of global clip index `clip` in a run seeded with `seed` (splitmix64 finaliser: distinct, well mixed)."""
    z = (int(seed) * 0x9E3779B97F4A7C15 + (int(clip) + 1) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return (z ^ (z >> 31)) & 0x7FFFFFFFFFFFFFFF


def chunk_steps_for(B, L, T, target_bytes=32 << 20):
    """Noise draws per staging chunk: ~32 MB of pinned memory per chunk, at least 1 draw."""
    return max(1, min(max(T - 1, 1), target_bytes // (4 * B * L)))


@torch.no_grad()
def sampling(net, size, diffusion_hyperparams, condition=None, use_graph=True, verbose=True, noise=None, out=None,
             chunk_steps=None):
    """Drop-in for generate.sampling(net, size, diffusion_hyperparams, condition).

    The reference draws x_T and then, inside the loop, one CPU normal tensor per step (generate.py:47,54).  Here the
    same draws are made from the same generator in the same order, but a chunk of steps ahead of the GPU: chunk k+1
    is drawn into pinned memory and copied on a side stream while the one-step CUDA graph replays chunk k
    (dwb_sample_steps), so the CPU generator never stalls the device.  `noise`: GlobalNoise() (default, the
    reference's process-wide generator) or PerClipNoise(seeds)."""
    dh = diffusion_hyperparams
    T = dh["T"]
    assert len(dh["Alpha"]) == T and len(dh["Alpha_bar"]) == T and len(dh["Sigma"]) == T and len(size) == 3
    if verbose:
        print("begin sampling, total number of reverse steps = %s" % T)
    eng = net._engine_get()
    noise = noise or GlobalNoise()
    B, _, L = size
    coef = step_coefficients(dh)
    k = chunk_steps or chunk_steps_for(B, L, T)
    sg = eng.staging(B, L, k)
    main = torch.cuda.current_stream(eng.device)
    side = sg["stream"]

    if sg["x_copied"] is not None:
        sg["x_copied"].synchronize()
    noise.fill(sg["x_host"])                                    # x_T: the first draw
    x = out if out is not None else torch.empty(tuple(size), device=eng.device)
    x.copy_(sg["x_host"], non_blocking=True)
    sg["x_copied"] = torch.cuda.Event()
    sg["x_copied"].record(main)

    t, c = T - 1, 0
    while t >= 0:
        n = min(k, t + 1)                                       # steps t, t-1, ..., t-n+1
        draws = n if t - n + 1 > 0 else n - 1                   # no draw at t = 0
        i = c & 1
        if draws > 0:
            if sg["copied"][i] is not None:
                sg["copied"][i].synchronize()                   # host chunk i is free again
            host = sg["host"][i]
            for j in range(draws):
                noise.fill(host[j])
            if sg["consumed"][i] is not None:
                side.wait_event(sg["consumed"][i])              # device chunk i is free again
            with torch.cuda.stream(side):
                sg["dev"][i][:draws].copy_(host[:draws], non_blocking=True)
                sg["copied"][i] = torch.cuda.Event()
                sg["copied"][i].record(side)
            main.wait_event(sg["copied"][i])
        eng.sample_steps(x, sg["dev"][i], coef, t, n, condition, use_graph=use_graph)
        sg["consumed"][i] = torch.cuda.Event()
        sg["consumed"][i].record(main)
        t -= n
        c += 1
    return x
