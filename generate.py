#!/usr/bin/env python
"""Generation entry point with the reference's command line (generate.py:203-231):

    python generate.py experiment=sc09 model=sashimi_small generate.n_samples=32 generate.batch_size=16
    python -m torch.distributed.run --nproc-per-node 8 generate.py ...        # one process per GPU

Differences from the reference are confined to what runs underneath: every diffusion step is a replay of one
CUDA graph inside libdwb with the CPU noise drawn a chunk of steps ahead, utterances of any length run on one
plan (kernel truncation below, overlap-save above the training segment length), the wav -> mel front end runs on
the GPU, and multi-GPU runs shard one globally defined batch with per-clip noise streams (so N-GPU output ==
1-GPU output) instead of spawning unsynchronised, unseeded processes.
`generate.random_init=true` writes samples from a seeded fresh model when no checkpoint exists
(useful on a box without weights; the reference would raise)."""
import os
import sys
import time

import torch

import diffwave_sashimi_b200 as dwb
from diffwave_sashimi_b200 import distributed as D
from diffwave_sashimi_b200 import experiment as E
from diffwave_sashimi_b200.config import compose

ROOT = os.path.dirname(os.path.abspath(__file__))


@torch.no_grad()
def generate(rank, world, diffusion_cfg, model_cfg, dataset_cfg, ckpt_iter="max", n_samples=1, name=None, batch_size=None,
             ckpt_smooth=None, mel_path=None, mel_name=None, seed=0, random_init=False, **_):
    torch.cuda.set_device(rank % torch.cuda.device_count())
    local_path, out_dir = E.local_directory(name, model_cfg, diffusion_cfg, dataset_cfg, "waveforms")
    dh = dwb.calc_diffusion_hyperparams(**{k: diffusion_cfg[k] for k in ("T", "beta_0", "beta_T", "beta")}, fast=True)
    net = dwb.construct_model(model_cfg)
    ckpt_dir = os.path.join("exp", local_path, "checkpoint")
    try:
        it, sd = E.load_state_dict(ckpt_dir, ckpt_iter, ckpt_smooth)
        net.load_state_dict(sd)
        print(f"Successfully loaded model at iteration {it}")
    except (FileNotFoundError, OSError):
        if not random_init:
            raise Exception("No valid model found")
        it = 0
        net.load_state_dict(dwb.init.seeded_state_dict(dict(model_cfg), seed=seed))
    net = net.cuda().eval()
    out_dir = os.path.join(out_dir, str(it))
    os.makedirs(out_dir, mode=0o775, exist_ok=True)

    cond = None
    if mel_name is not None:
        if mel_path is not None:       # pre-generated spectrogram (generate.py:136-142 of the reference)
            cond = torch.load(os.path.join(mel_path, f"{mel_name}.wav.pt")).unsqueeze(0).cuda()
        else:                          # wav -> mel on the GPU (reference: dataloaders.mel2samp.Mel2Samp.get_mel, :143-155)
            from diffwave_sashimi_b200 import mel as M
            stft = M.TacotronSTFT(**{k: dataset_cfg[k] for k in ("filter_length", "hop_length", "win_length", "sampling_rate",
                                                                "mel_fmin", "mel_fmax")})
            audio, sr = M.load_wav_to_torch(os.path.join(dataset_cfg["data_path"], f"{mel_name}.wav"))
            if sr != dataset_cfg["sampling_rate"]:
                raise ValueError(f"{sr} SR doesn't match target {dataset_cfg['sampling_rate']} SR")
            cond = M.get_mel(stft, audio.cuda()).unsqueeze(0)
        audio_length = cond.shape[-1] * dataset_cfg["hop_length"]
    else:
        audio_length = dataset_cfg["segment_length"]
    total = n_samples * world                     # n_samples is per GPU, like the reference
    bs = (batch_size or n_samples) * world
    assert total % bs == 0
    print(f"begin generating audio of length {audio_length} | {total} samples with global batch size {bs}")
    t0 = time.time()
    chunks = [D.generate_sharded(net, bs, audio_length, dh, seed + i, cond, rank, world) for i in range(total // bs)]
    audio = torch.cat(chunks)
    torch.cuda.synchronize()
    print(f"generated {total} samples shape {tuple(audio.shape)} at iteration {it} in {time.time() - t0:.2f} seconds")
    if rank == 0:
        from scipy.io.wavfile import write as wavwrite
        for i in range(total):
            wavwrite(os.path.join(out_dir, f"{it // 1000}k_{i}.wav"), dataset_cfg["sampling_rate"],
                     audio[i].squeeze().cpu().numpy())
        print(f"saved generated samples at iteration {it}")
    return audio


def main(argv=None):
    cfg = compose(os.path.join(ROOT, "configs"), "config", list(sys.argv[1:] if argv is None else argv))
    rank, world = D.init()
    generate(rank, world, cfg.diffusion, cfg.model, cfg.dataset, **cfg.generate)


if __name__ == "__main__":
    main()
