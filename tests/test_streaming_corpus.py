"""Streaming (config 3) and corpus sharding (config 5): host logic on CPU (gloo, world_size 2) and
GPU equivalence of chunked streaming with the one-shot transform."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import rel_err


def test_chunk_plan_covers_every_frame_once():
    from tal_asrd_b200.streaming import chunk_plan
    for L, cf in ((57_600_000, 3000), (16000, 7), (201, 3000), (480_123, 1000), (16000, 1), (3200, 2)):
        plan = chunk_plan(L, cf)
        T = 1 + L // 160
        assert plan[0][0] == 0 and plan[-1][1] == T
        for (a0, a1, lo, hi), (b0, _, _, _) in zip(plan, plan[1:]):
            assert a1 == b0
        for f0, f1, lo, hi in plan:
            assert 0 <= lo < hi <= L
            # every source sample of every frame of the chunk (after reflection) is inside [lo, hi)
            g = np.arange(160 * f0 - 200, 160 * (f1 - 1) + 200)
            g = np.where(g < 0, -g, g)
            g = np.where(g >= L, 2 * (L - 1) - g, g)
            assert g.min() >= lo and g.max() < hi
            assert hi - lo <= 160 * (f1 - f0) + 240 + 200


def test_shard_episodes_partitions_and_balances():
    from tal_asrd_b200.corpus import shard_episodes
    rng = np.random.default_rng(0)
    lengths = rng.integers(16000, 9_600_000, size=97).tolist()
    for world in (1, 2, 4, 8):
        shards = [shard_episodes(lengths, world, r) for r in range(world)]
        assert sorted(sum(shards, [])) == list(range(97))
        loads = [sum(lengths[i] for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(lengths)
    with pytest.raises(ValueError):
        shard_episodes(lengths, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, lengths, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tal_asrd_b200.corpus import CorpusStats, shard_episodes
    mine = shard_episodes(lengths, world, rank)
    stats = CorpusStats(n_mels=4)
    for ep in mine:                                   # stand-in feature blocks: deterministic per episode
        g = torch.Generator().manual_seed(ep)
        feats = torch.randn(lengths[ep] // 1000, 4, generator=g, dtype=torch.float64) * 2 + 1
        block = torch.zeros(1, 11, dtype=torch.float64)
        block[0, 0] = feats.numel()
        block[0, 1] = feats.sum()
        block[0, 2] = (feats ** 2).sum()
        block[0, 3:7] = feats.sum(0)
        block[0, 7:11] = (feats ** 2).sum(0)
        stats.add(block)
    stats.all_reduce()
    q.put((rank, mine, stats.block.clone().numpy(), stats.mean, stats.var, stats.mel_mean.numpy(), stats.mel_var.numpy()))
    dist.destroy_process_group()


def test_corpus_stats_allreduce_gloo_world2():
    lengths = [20000, 5000, 9000, 14000, 3000, 11000, 7000]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, lengths, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res[0][1] + res[1][1]) == list(range(len(lengths)))       # episodes partitioned
    assert np.array_equal(res[0][2], res[1][2])                              # both ranks hold the global sums
    allf = []
    for ep in range(len(lengths)):
        g = torch.Generator().manual_seed(ep)
        allf.append(torch.randn(lengths[ep] // 1000, 4, generator=g, dtype=torch.float64) * 2 + 1)
    allf = torch.cat(allf).numpy()
    assert abs(res[0][3] - allf.mean()) < 1e-12 and abs(res[0][4] - allf.var()) < 1e-10
    assert np.allclose(res[0][5], allf.mean(0), atol=1e-12) and np.allclose(res[0][6], allf.var(0), atol=1e-10)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_streamed_episode_equals_one_shot():
    """Chunked streaming (host -> device, 30 s chunks with 200-sample halos) must give the one-shot grid."""
    from oracle import logmel_oracle as O
    from tal_asrd_b200 import LogMelSpec, synth
    from tal_asrd_b200.streaming import stream_episode
    dev = torch.device("cuda:0")
    mod = LogMelSpec().to(dev)
    L = 16000 * 200 + 77                                                     # 200 s, ragged tail
    ep = torch.from_numpy(synth.waveform(2020, 9, 0, L))
    one = mod(ep[None].to(dev))
    for chunk_s in (30.0, 7.3, 0.5):
        got = stream_episode(mod, ep.pin_memory(), chunk_seconds=chunk_s, device=dev)
        torch.cuda.synchronize()
        assert got.shape == one.shape
        assert float((got - one).abs().max()) < 2e-5, chunk_s            # only the summation order of the mean differs
    got = stream_episode(mod, ep.to(dev), chunk_seconds=11.0, coalesce_on_device=False)               # device-resident episode
    assert float((got - one).abs().max()) < 2e-5
    # device-resident episodes have nothing to stage: by default their chunks are coalesced (here: one launch); the
    # un-normalised features do not depend on the chunking at all
    raw_small = stream_episode(mod, ep.to(dev), chunk_seconds=11.0, normalise=False, coalesce_on_device=False)
    raw_coalesced = stream_episode(mod, ep.to(dev), chunk_seconds=11.0, normalise=False)
    assert torch.equal(raw_small, raw_coalesced)
    assert torch.equal(stream_episode(mod, ep.to(dev), chunk_seconds=11.0), one)
    ref = O.logmel_f64(ep[None, :160 * 300].numpy())                         # oracle on a prefix (interior frames agree up to the mean)
    d = got[0, :250].cpu().numpy() - ref[0, :250]
    assert np.abs(d - d.mean()).max() < 1e-4
    # int16 PCM streamed
    pcm = torch.from_numpy(synth.pcm16(2020, 9, 0, L))
    got16 = stream_episode(mod, pcm, chunk_seconds=30.0, device=dev)
    assert float((got16 - one).abs().max()) < 2e-5


@pytest.mark.gpu
def test_hour_long_episode_streams():
    """BASELINE config 3 at full size: 1 h = 57.6 M samples -> 360 001 frames, streamed in 30 s chunks."""
    from tal_asrd_b200 import LogMelSpec, _lib
    from tal_asrd_b200.streaming import stream_episode
    dev = torch.device("cuda:0")
    mod = LogMelSpec().to(dev)
    L = 57_600_000
    ep = torch.empty(L, device=dev)
    _lib.check(_lib.load().talfe_synth_fill(ep.data_ptr(), _lib.F32, 1, L, L, 2020, 1234, 0, None))
    one = mod(ep[None])
    got = stream_episode(mod, ep, chunk_seconds=30.0, coalesce_on_device=False)
    torch.cuda.synchronize()
    assert got.shape == (1, 360001, 80)
    assert float((got - one).abs().max()) < 2e-5
    assert abs(float(got.double().mean())) < 2e-6


@pytest.mark.gpu
def test_corpus_pass_single_rank_global_cmvn():
    from oracle import logmel_oracle as O
    from tal_asrd_b200 import LogMelSpec, synth
    from tal_asrd_b200.corpus import corpus_pass
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    mod = LogMelSpec().to(dev)
    lens = [48000, 100000, 16001]
    eps_ = [torch.from_numpy(synth.waveform(1, i, 0, n)) for i, n in enumerate(lens)]
    feats, stats = corpus_pass(mod, eps_, norm="row_mel_var", chunk_seconds=2.0)
    torch.cuda.synchronize()
    raw = np.concatenate([O.logmel_unnormalised_f64(e[None].numpy())[0] for e in eps_])
    assert stats.count == raw.size
    assert np.allclose(stats.mel_mean.cpu().numpy(), raw.mean(0), atol=1e-5)
    assert np.allclose(stats.mel_var.cpu().numpy(), raw.var(0), rtol=1e-4)
    got = np.concatenate([f[0].cpu().numpy() for f in feats])
    want = (raw - raw.mean(0)) / raw.std(0)
    assert rel_err(got, want) < 5e-4
    # pass 2 for corpora whose features do not stay resident: the global statistics applied inside the transform kernel
    from tal_asrd_b200.corpus import corpus_second_pass
    again = np.concatenate([f[0].cpu().numpy() for f in corpus_second_pass(mod, eps_, stats)])
    assert rel_err(again, want) < 5e-4
    assert np.abs(again - got).max() < 2e-5          # (streamed 2-s chunks vs one shot: chunk-edge tiles regroup the sums)


def _nccl_corpus_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from tal_asrd_b200 import LogMelSpec, synth
    from tal_asrd_b200.corpus import corpus_pass, shard_episodes
    lengths = [64000, 48000, 100000, 16001, 80000]
    mine = shard_episodes(lengths, world, rank)
    mod = LogMelSpec().to(torch.device("cuda", rank))
    eps_ = [torch.from_numpy(synth.waveform(1, i, 0, lengths[i])) for i in mine]
    feats, stats = corpus_pass(mod, eps_, norm="row_mel_var", chunk_seconds=2.0)
    torch.cuda.synchronize()
    q.put((rank, mine, stats.block.cpu().numpy(), [f[0].cpu().numpy() for f in feats]))
    dist.destroy_process_group()


@pytest.mark.gpu
def test_corpus_pass_two_ranks_nccl_allreduce():
    """BASELINE config 5 in miniature: episodes sharded over 2 GPUs, ONE all-reduce of the statistics block."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from oracle import logmel_oracle as O
    from tal_asrd_b200 import synth
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_corpus_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=90) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    lengths = [64000, 48000, 100000, 16001, 80000]
    raw = {i: O.logmel_unnormalised_f64(synth.waveform(1, i, 0, n)[None])[0] for i, n in enumerate(lengths)}
    allraw = np.concatenate([raw[i] for i in range(len(lengths))])
    assert np.array_equal(res[0][2], res[1][2])                            # both ranks hold the global sums
    assert res[0][2][0, 0] == allraw.size
    mu, sd = allraw.mean(0), allraw.std(0)
    for rank, mine, _, feats in res:
        for i, f in zip(mine, feats):
            assert rel_err(f, (raw[i] - mu) / sd) < 5e-4


def _nccl_c_abi_worker(rank, world, port, q):
    """talfe_allreduce_stats with an ncclComm_t created through NCCL's own C API (no torch.distributed collectives)."""
    import ctypes
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)            # only to ship the 128-byte unique id
    import glob
    import nvidia.nccl  # noqa: F401  (torch's bundled NCCL; a namespace package: use __path__)
    cands = sorted(glob.glob(os.path.join(list(nvidia.nccl.__path__)[0], "lib", "libnccl.so*")))
    nccl = ctypes.CDLL(cands[0], mode=ctypes.RTLD_GLOBAL)
    uid = (ctypes.c_byte * 128)()
    if rank == 0:
        assert nccl.ncclGetUniqueId(ctypes.byref(uid)) == 0
    t = torch.tensor(list(bytes(uid)), dtype=torch.uint8)
    dist.broadcast(t, 0)
    uid = (ctypes.c_byte * 128).from_buffer_copy(bytes(t.tolist()))
    comm = ctypes.c_void_p()

    class UID(ctypes.Structure):
        _fields_ = [("internal", ctypes.c_byte * 128)]
    nccl.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, UID, ctypes.c_int]
    u = UID()
    ctypes.memmove(ctypes.byref(u), uid, 128)
    assert nccl.ncclCommInitRank(ctypes.byref(comm), world, u, rank) == 0
    from tal_asrd_b200 import _lib
    lib = _lib.load()
    stats = torch.arange(163, dtype=torch.float64, device=f"cuda:{rank}") * (rank + 1)
    rc = lib.talfe_allreduce_stats(stats.data_ptr(), 163, comm, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    q.put((rank, rc, stats.cpu().numpy()))
    nccl.ncclCommDestroy.argtypes = [ctypes.c_void_p]
    nccl.ncclCommDestroy(comm)
    dist.destroy_process_group()


@pytest.mark.gpu
def test_c_abi_allreduce_stats_two_ranks():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_c_abi_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=90) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.arange(163, dtype=np.float64) * 3
    for rank, rc, got in res:
        assert rc == 0 and np.array_equal(got, want)


def test_host_pipeline_refuses_to_run_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from tal_asrd_b200 import HostPipeline, LogMelSpec
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        HostPipeline(LogMelSpec())


@pytest.mark.gpu
def test_host_pipeline_equals_forward_per_batch():
    """Seven different host batches through two device slots (every slot reused three times, shapes and dtypes
    changing on the way): each result must be bit-identical to LogMelSpec.forward on that batch alone, i.e. no
    batch may see a slot before the previous occupant has been transformed / copied out."""
    from oracle import logmel_oracle as O
    from tal_asrd_b200 import HostPipeline, LogMelSpec, synth
    dev = torch.device("cuda:0")
    mod = LogMelSpec().to(dev)
    pipe = HostPipeline(mod, dev, depth=2)
    shapes = [(6, 160000)] * 4 + [(3, 48000 + 77)] * 2 + [(6, 160000)]
    ins, outs, evs = [], [], []
    for i, (b, n) in enumerate(shapes):
        x = torch.from_numpy(synth.batch(31 + i, b, n))
        if i == 5:
            x = (x * 32768.0).round().clamp_(-32768, 32767).to(torch.int16)
        ins.append(x.pin_memory())
        outs.append(torch.full((b, 1 + n // 160, 80), float("nan")).pin_memory())
        evs.append(pipe.submit(ins[-1], outs[-1]))
    pipe.drain()
    assert all(e.query() for e in evs)
    for i, (x, y) in enumerate(zip(ins, outs)):
        want = mod(x.to(dev)).cpu()
        assert torch.equal(y, want), f"batch {i}"
    ref = O.logmel_f64(ins[0].numpy())
    assert rel_err(outs[0].numpy(), ref) < 1e-4
    with pytest.raises(ValueError):
        pipe.submit(ins[0], outs[1][:, :10])
    with pytest.raises(ValueError):
        pipe.submit(ins[0].to(dev), outs[0])
    # per-row semantics and the [B, 80, T] layout ride through unchanged
    pipe2 = HostPipeline(mod, dev, depth=1, norm="row", layout="mt")
    lens = torch.tensor([160000, 16000, 201, 99999, 160000, 7777])
    y2 = torch.empty(6, 80, 1001).pin_memory()
    pipe2.submit(ins[0], y2, audio_lens=lens).synchronize()
    want = mod.features(ins[0].to(dev), audio_lens=lens, norm="row", layout="mt").cpu()
    assert torch.equal(y2, want)
