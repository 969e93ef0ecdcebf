"""diffwave_sashimi_b200 — B200-native engine behind the plugin surface of albertfgu/diffwave-sashimi.

Public API (mirrors the reference's names):
    construct_model(model_cfg), model_identifier(model_cfg)       models/__init__.py:4-23
    WaveNet, Sashimi                                              models/wavenet.py, models/sashimi.py
    sampling(net, size, diffusion_hyperparams, condition=None)    generate.py:23-55
    calc_diffusion_hyperparams(T, beta_0, beta_T, beta, fast)     utils.py:121-151
    ops.cauchy_mult(v, z, w, symmetric=True)                      extensions/cauchy/cauchy.py:46-63
    training.Trainer(net, batch, length, lr)                      train.py:84-143,198-222 (loss + backward + Adam, wavenet)
"""
from . import init, ops, training  # noqa: F401
from .engine import Engine  # noqa: F401
from .models import Sashimi, WaveNet, construct_model, model_identifier  # noqa: F401
from .sampler import (GlobalNoise, PerClipNoise, calc_diffusion_hyperparams, clip_seed, draw_noise,  # noqa: F401
                      sampling, step_coefficients)
