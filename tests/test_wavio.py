"""CPU: the WAV -> pinned int16 loader against the reference loader's slice arithmetic
(tal/asr/data/util.py:32-43) and against torchaudio-style normalisation."""
import os
import struct

import numpy as np
import pytest
import torch

from tal_asrd_b200 import synth, wavio


@pytest.fixture()
def wav(tmp_path):
    pcm = synth.pcm16(7, 3, 0, 16000 * 12 + 37)
    path = str(tmp_path / "ep.wav")
    wavio.write_wav_pcm16(path, pcm)
    return path, pcm


def test_info_and_whole_file(wav):
    path, pcm = wav
    info = wavio.wav_info(path)
    assert (info.rate, info.channels, info.length, info.data_offset) == (16000, 1, pcm.size, 44)
    ep = wavio.load_episode_pcm16(path)
    assert ep.dtype == torch.int16 and ep.dim() == 1 and np.array_equal(ep.numpy(), pcm)


@pytest.mark.parametrize("start_s,end_s", [(0.0, 1.0), (2.5, 7.25), (11.0, None), (3.3333, 3.9999), (0.0, 100.0)])
def test_segment_matches_reference_slice_arithmetic(wav, start_s, end_s):
    """offset = int(start_s * rate); num_frames = int((end_s - start_s) * rate), 0 = to the end (util.py:35-41)."""
    path, pcm = wav
    seg = wavio.load_audio_segment_pcm16(path, start_s, end_s)
    offset = int(start_s * 16000)
    frames = int((end_s - start_s) * 16000) if end_s is not None else 0
    want = pcm[offset:] if frames <= 0 else pcm[offset:offset + frames]
    assert np.array_equal(seg.numpy(), want)
    # what the reference hands the front end: the same samples / 32768 as float32 (util.py:43)
    assert np.array_equal(seg.float().numpy() / 32768.0, want.astype(np.float32) / np.float32(32768.0))


def test_matches_torchaudio_when_its_backend_is_available(wav):
    path, pcm = wav
    torchaudio = pytest.importorskip("torchaudio")
    try:
        x, sr = torchaudio.load(path, frame_offset=16000, num_frames=32000)
    except Exception as exc:                                             # no I/O backend in this image
        pytest.skip(f"torchaudio.load unavailable: {exc}")
    seg = wavio.load_audio_segment_pcm16(path, 1.0, 3.0)
    assert sr == 16000 and torch.equal(x[0], seg.float() / 32768.0)


def test_collate_layout_is_the_collaters(wav):
    """zero right-pad to the longest + audio_lens (tal/asr/data/aligned.py:246-270)."""
    path, pcm = wav
    segs = wavio.segments_from_files([(path, 0.0, 1.0), (path, 2.0, 5.5), (path, 10.0, None)])
    batch, lens = wavio.collate_pcm16(segs)
    assert batch.dtype == torch.int16 and tuple(batch.shape) == (3, max(s.numel() for s in segs))
    assert lens.tolist() == [s.numel() for s in segs]
    for b, s in enumerate(segs):
        assert torch.equal(batch[b, :s.numel()], s) and not batch[b, s.numel():].any()


def test_out_buffer_and_errors(wav, tmp_path):
    path, pcm = wav
    out = torch.empty(20000, dtype=torch.int16)
    seg = wavio.load_audio_segment_pcm16(path, 0.5, 1.5, out=out)
    assert seg.data_ptr() == out.data_ptr() and seg.numel() == 16000
    with pytest.raises(ValueError):
        wavio.load_audio_segment_pcm16(path, 0.0, 2.0, out=torch.empty(10, dtype=torch.int16))
    with pytest.raises(ValueError):
        wavio.load_audio_segment_pcm16(path, 100.0, 101.0)                # offset beyond the file
    bad = str(tmp_path / "rate.wav")
    wavio.write_wav_pcm16(bad, pcm[:100], rate=8000)
    with pytest.raises(ValueError):
        wavio.load_audio_segment_pcm16(bad)                                # the reference would resample; this loader refuses
    f32 = str(tmp_path / "f32.wav")
    with open(f32, "wb") as fh:                                            # IEEE float WAV: format tag 3
        data = np.zeros(10, np.float32).tobytes()
        fh.write(b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE" + b"fmt " +
                 struct.pack("<IHHIIHH", 16, 3, 1, 16000, 64000, 4, 32) + b"data" + struct.pack("<I", len(data)) + data)
    with pytest.raises(ValueError):
        wavio.wav_info(f32)
    notwav = str(tmp_path / "x.bin")
    open(notwav, "wb").write(b"hello world, not a wav")
    with pytest.raises(ValueError):
        wavio.wav_info(notwav)


def test_header_with_extra_chunks(tmp_path):
    pcm = synth.pcm16(1, 1, 0, 999)
    path = str(tmp_path / "list.wav")
    with open(path, "wb") as fh:
        junk = b"LIST" + struct.pack("<I", 5) + b"abcde" + b"\x00"        # odd-sized chunk + pad byte
        fh.write(b"RIFF" + struct.pack("<I", 36 + len(junk) + pcm.nbytes) + b"WAVE")
        fh.write(b"fmt " + struct.pack("<IHHIIHH", 16, 1, 1, 16000, 32000, 2, 16))
        fh.write(junk)
        fh.write(b"data" + struct.pack("<I", pcm.nbytes) + pcm.tobytes())
    assert np.array_equal(wavio.load_episode_pcm16(path).numpy(), pcm)
