"""B200-native (sm_100a) spectral losses: drop-in for the reference's losses/stft_loss.py and
losses/mel_loss.py (s194584/dl-speech-enhancement), backed by hand-written CUDA in libspecloss.so."""
from .modules import (LogSTFTMagnitudeLoss, MelSpectrogram, MultiWindowShapeLoss, WaveformShapeLoss, MultiMelSpectrogramLoss, MultiResolutionSTFTLoss,
                      SpectralConvergenceLoss, SpectralLoss, STFTLoss, spectrogram, stft, MelL1, Mel_L1)
from .functional import spectral_losses

__all__ = ["stft", "spectrogram", "SpectralConvergenceLoss", "LogSTFTMagnitudeLoss", "STFTLoss", "MultiResolutionSTFTLoss",
           "MelSpectrogram", "MultiMelSpectrogramLoss", "SpectralLoss", "spectral_losses",
           "WaveformShapeLoss", "MultiWindowShapeLoss", "MelL1", "Mel_L1"]
