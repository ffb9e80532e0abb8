"""Sharded ciphertext sweep (BASELINE config 5): host-side partition logic, the
world-size-2 `gloo` run on CPU (the CTA emulator stands in for the device) and the single-GPU run.
The oracle is only the checker."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

LOGN, BITS, PBITS, TOTAL, WAVE = 10, [40, 30], 40, 5, 2


def test_shard_range_is_a_contiguous_balanced_partition():
    from hehub_b200.sweep import shard_range
    for total in (0, 1, 5, 8, 65536, 65537):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f1 == f0 + c0
            assert spans[-1][0] + spans[-1][1] == total
            counts = [c for _, c in spans]
            assert max(counts) - min(counts) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_shard_range_of_the_driver_matches_the_python_restatement():
    import ctypes as C
    import __graft_entry__ as ge
    from hehub_b200.binding import load_library
    from hehub_b200.sweep import shard_range
    lib = load_library(ge.build_sim())
    lib.hehub_b200_shard_range.argtypes = [C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    lib.hehub_b200_shard_range.restype = None
    for total in (0, 1, 5, 8, 65536, 65537):
        for world in (1, 2, 3, 8):
            for rank in range(world):
                f, n = C.c_size_t(), C.c_size_t()
                lib.hehub_b200_shard_range(total, world, rank, C.byref(f), C.byref(n))
                assert (f.value, n.value) == shard_range(total, world, rank)


@pytest.mark.parametrize("kind", ["sim", pytest.param("gpu", marks=pytest.mark.gpu)])
def test_checksum_kernel_matches_numpy(kind):
    import __graft_entry__ as ge
    from hehub_b200.binding import Context
    from hehub_b200.sweep import ct_checksum_numpy
    ctx = Context(lib_path=ge.build_sim()) if kind == "sim" else Context(device=0)
    try:
        rng = np.random.default_rng(1)
        for cts, words in ((3, 64), (1, 1), (5, 4099), (2, 98304)):
            w = rng.integers(0, 1 << 63, (cts, words), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, (cts, words), dtype=np.uint64)
            d, sums = ctx.to_device(w), ctx.slab(cts)
            ctx._call("ct_checksums", d.ptr, words, cts, sums.ptr)
            got = sums.download()
            assert [int(g) for g in got] == [ct_checksum_numpy(w[i]) for i in range(cts)]
            d.free()
            sums.free()
    finally:
        ctx.close()


def _moduli(oracle):
    mods, p = oracle.ckks_pick_moduli(BITS, PBITS)
    return [int(m) for m in mods], int(p)


def _check_sample(oracle, sweep, checksums, indices):
    from hehub_b200.sweep import ct_checksum_numpy
    key = sweep.key_host()
    for idx in indices:
        ct1, ct2 = sweep.one_ct_inputs_host(idx)
        want = oracle.ckks_mult_relin(sweep.logn, sweep.ext, ct1, ct2, key)
        assert ct_checksum_numpy(want) == int(checksums[idx]), f"ciphertext {idx}"


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


WORKER = r"""
import ctypes as C, os, sys, json
import numpy as np
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
import __graft_entry__ as ge
from hehub_b200.binding import Context
from hehub_b200.sweep import CtSweep, callback_collectives
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)

def view(address, nbytes):  # the emulator's "device" memory is host memory
    return torch.frombuffer((C.c_uint8 * nbytes).from_address(address), dtype=torch.uint8)

def broadcast(address, nbytes, root):
    dist.broadcast(view(address, nbytes), src=root)

def allgather(send, recv, nbytes):
    out = view(recv, nbytes * world)
    dist.all_gather_into_tensor(out, view(send, nbytes).clone()) if hasattr(dist, "all_gather_into_tensor") and False else \
        dist.all_gather(list(out.view(world, nbytes).unbind(0)), view(send, nbytes).clone())

ctx = Context(lib_path=ge.SIM_SO)            # CTA emulator build of the kernel sources (tests only)
sw = CtSweep(ctx, {logn}, {mods!r}, {p}, seed=7, rank=rank, world=world, collectives=callback_collectives(broadcast, allgather))
res = sw.run({total}, {wave})
if rank == 0:
    print("RESULT " + json.dumps({{"all": [int(v) for v in res["all_checksums"]], "count": res["count"]}}), flush=True)
else:
    assert "all_checksums" not in res
sw.close()
dist.destroy_process_group()
ctx.close()
"""


def test_world_size_2_gloo_sweep_matches_single_process_and_oracle(oracle):
    import __graft_entry__ as ge
    from hehub_b200.binding import Context
    from hehub_b200.sweep import CtSweep
    ge.build_sim()
    mods, p = _moduli(oracle)
    # single process
    ctx = Context(lib_path=ge.SIM_SO)
    try:
        sw = CtSweep(ctx, LOGN, mods, p, seed=7)
        single = sw.run(TOTAL, WAVE)
        assert single["count"] == TOTAL
        _check_sample(oracle, sw, single["checksums"], [0, TOTAL - 1])
    finally:
        ctx.close()
    # two ranks over gloo
    port = _free_port()
    code = WORKER.format(root=ROOT, logn=LOGN, mods=mods, p=p, total=TOTAL, wave=WAVE)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [pr.communicate(timeout=600) for pr in procs]
    for pr, (so, se) in zip(procs, outs):
        assert pr.returncode == 0, se[-2000:]
    import json
    line = [l for l in outs[0][0].splitlines() if l.startswith("RESULT ")][0]
    got = json.loads(line[len("RESULT "):])
    assert got["count"] == 3  # ranks take 3 + 2 of the 5 ciphertexts
    assert got["all"] == [int(v) for v in single["checksums"]]


@pytest.mark.gpu
def test_single_gpu_sweep_matches_oracle(oracle):
    from hehub_b200.binding import Context
    from hehub_b200.sweep import CtSweep
    mods, p = oracle.ckks_pick_moduli([40, 30, 30, 30], 40)
    ctx = Context(device=0)
    try:
        sw = CtSweep(ctx, 12, [int(m) for m in mods], int(p), seed=11)
        res = sw.run(11, 4, time_ops=True)
        assert res["count"] == 11 and len(set(int(v) for v in res["checksums"])) == 11 and res["op_seconds"] > 0
        _check_sample(oracle, sw, res["checksums"], [0, 5, 10])
        sw.close()
    finally:
        ctx.close()


NCCL_WORKER = r"""
import os, sys, json
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
from hehub_b200.binding import Context
from hehub_b200.sweep import CtSweep, nccl_collectives
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("gloo", rank=rank, world_size=world)   # only carries the 128-byte NCCL id

def exchange(raw):
    box = [raw]
    dist.broadcast_object_list(box, src=0)
    return box[0]

coll, destroy = nccl_collectives(rank, world, rank, exchange)
ctx = Context(device=rank)
sw = CtSweep(ctx, {logn}, {mods!r}, {p}, seed=7, rank=rank, world=world, collectives=coll)
res = sw.run({total}, {wave})
if rank == 0:
    print("RESULT " + json.dumps({{"all": [int(v) for v in res["all_checksums"]]}}), flush=True)
sw.close(); destroy(); ctx.close()
dist.destroy_process_group()
"""


@pytest.mark.gpu
def test_two_gpu_nccl_sweep_matches_single_gpu(oracle):
    """The C++ driver with the NCCL provider on two GPUs of one box: same checksums as one GPU, pair for pair."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("one GPU on this box")
    from hehub_b200.binding import Context
    from hehub_b200.sweep import CtSweep
    mods, p = _moduli(oracle)
    ctx = Context(device=0)
    try:
        sw = CtSweep(ctx, LOGN, mods, p, seed=7)
        single = sw.run(TOTAL, WAVE)
        sw.close()
    finally:
        ctx.close()
    port = _free_port()
    code = NCCL_WORKER.format(root=ROOT, logn=LOGN, mods=mods, p=p, total=TOTAL, wave=WAVE)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [pr.communicate(timeout=600) for pr in procs]
    for pr, (so, se) in zip(procs, outs):
        assert pr.returncode == 0, se[-2000:]
    assert "NCCL" in outs[0][1] and "communicator" in outs[0][1]
    import json
    line = [l for l in outs[0][0].splitlines() if l.startswith("RESULT ")][0]
    assert json.loads(line[len("RESULT "):])["all"] == [int(v) for v in single["checksums"]]
