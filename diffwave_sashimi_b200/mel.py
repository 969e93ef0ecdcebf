"""Host mirror of the reference's mel front end (dataloaders/stft.py `TacotronSTFT`, dataloaders/mel2samp.py
`Mel2Samp.get_mel`) on libdwb: the tables are built here exactly the way the reference builds them (numpy fft of
the identity, scipy window), the arithmetic runs in `dwb_mel_spectrogram`.  librosa is not needed: its two helpers
on this path are restated below (`librosa.util.pad_center`, `librosa.filters.mel` with the Slaney scale and norm)."""
import ctypes

import numpy as np
import torch

from ._lib import check, lib, ptr, stream_ptr

MAX_WAV_VALUE = 32768.0          # mel2samp.py:17


def _hz_to_mel(f):
    """Slaney mel scale (librosa.hz_to_mel, htk=False): linear below 1 kHz, log above."""
    f = np.asarray(f, dtype=np.float64)
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep, f / f_sp)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr, n_fft, n_mels=80, fmin=0.0, fmax=None):
    """librosa.filters.mel(sr=, n_fft=, n_mels=, fmin=, fmax=) with its defaults htk=False, norm='slaney':
    triangular filters on the Slaney mel scale, each scaled to unit area.  (n_mels, n_fft//2 + 1) float32."""
    fmax = sr / 2.0 if fmax is None else fmax
    fftfreqs = np.linspace(0, sr / 2.0, n_fft // 2 + 1)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    w = np.zeros((n_mels, n_fft // 2 + 1))
    for i in range(n_mels):
        w[i] = np.maximum(0, np.minimum(-ramps[i] / fdiff[i], ramps[i + 2] / fdiff[i + 1]))
    w *= (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]
    return w.astype(np.float32)


def forward_basis(filter_length, win_length, window="hann"):
    """`STFT.forward_basis` (stft.py:110-133) as (2*(n/2+1), n) float32: real then imaginary rows of the DFT matrix,
    times the periodic window zero-padded symmetrically to filter_length."""
    from scipy.signal import get_window
    fb = np.fft.fft(np.eye(filter_length))
    cutoff = filter_length // 2 + 1
    fb = np.vstack([np.real(fb[:cutoff, :]), np.imag(fb[:cutoff, :])])
    basis = torch.FloatTensor(fb)
    if window is not None:
        assert filter_length >= win_length
        win = get_window(window, win_length, fftbins=True)
        lpad = (filter_length - win_length) // 2                      # librosa.util.pad_center
        win = np.pad(win, (lpad, filter_length - win_length - lpad))
        basis = basis * torch.from_numpy(win).float()
    return basis.float()


class TacotronSTFT:
    """Same constructor and `mel_spectrogram(y)` as the reference class (stft.py:197-244); y (B, T) in [-1, 1] on a
    CUDA device -> (B, n_mel_channels, T // hop + 1)."""

    def __init__(self, filter_length=1024, hop_length=256, win_length=1024, n_mel_channels=80, sampling_rate=22050,
                 mel_fmin=0.0, mel_fmax=8000.0):
        self.filter_length, self.hop_length, self.win_length = filter_length, hop_length, win_length
        self.n_mel_channels, self.sampling_rate = n_mel_channels, sampling_rate
        self.forward_basis = forward_basis(filter_length, win_length)                       # (2 nb, n)
        self.mel_basis = torch.from_numpy(mel_filterbank(sampling_rate, filter_length, n_mel_channels, mel_fmin, mel_fmax))
        self._dev = {}

    def _tables(self, device):
        if device not in self._dev:
            self._dev[device] = (self.forward_basis.t().contiguous().to(device), self.mel_basis.contiguous().to(device))
        return self._dev[device]

    @torch.no_grad()
    def mel_spectrogram(self, y, in_scale=1.0):
        if not y.is_cuda:
            raise RuntimeError("the mel front end runs on the GPU (libdwb); there is no CPU path")
        y = y.to(torch.float32).contiguous()
        B, T = y.shape
        basis_t, melb = self._tables(y.device)
        frames = ctypes.c_int(0)
        check(lib().dwb_mel_frames(T, self.filter_length, self.hop_length, ctypes.byref(frames)))
        out = torch.empty(B, self.n_mel_channels, frames.value, dtype=torch.float32, device=y.device)
        with torch.cuda.device(y.device):
            check(lib().dwb_mel_spectrogram(ptr(y), B, T, float(in_scale), ptr(basis_t), self.filter_length, self.hop_length,
                                            ptr(melb), self.n_mel_channels, 1e-5, ptr(out), stream_ptr(y.device)))
        return out


def get_mel(stft: TacotronSTFT, audio):
    """`Mel2Samp.get_mel` (mel2samp.py:78-84): int16-valued wav samples (T,) -> (n_mels, frames)."""
    return stft.mel_spectrogram(audio.reshape(1, -1), in_scale=1.0 / MAX_WAV_VALUE)[0]


def load_wav_to_torch(path):
    """mel2samp.py:27-32"""
    from scipy.io.wavfile import read
    sr, data = read(path)
    return torch.from_numpy(np.asarray(data)).float(), sr
