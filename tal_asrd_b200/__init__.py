"""B200-native log-mel acoustic front end: drop-in for calclavia/tal-asrd's ``LogMelSpec``.

Public surface:
    LogMelSpec            nn.Module with the reference's constructor / forward contract
    num_frames            1 + L // 160 (RuntimeError for L <= 200)
    reference_tables      fp32 window / mel filterbank, bit-identical to the reference's buffers
    HostPipeline          host-resident batches with H2D | transform | D2H overlapped across batches
    Resample              torchaudio.transforms.Resample(orig_freq, new_freq) for device-resident waveforms (loader side)
    synth                 deterministic synthetic audio (numpy; CUDA twin: talfe_synth_fill)
"""
from . import specaug, synth  # noqa: F401
from .frontend import DEFAULT_SR, LogMelSpec, num_frames, reference_tables  # noqa: F401
from .hostpipe import HostPipeline  # noqa: F401
from .resample import Resample  # noqa: F401

__all__ = ["LogMelSpec", "num_frames", "reference_tables", "synth", "DEFAULT_SR", "HostPipeline", "Resample"]
