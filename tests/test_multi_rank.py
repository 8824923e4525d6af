"""N > 1 host logic on CPU: world_size 2 over gloo.  Each rank analyses its utterance shard (the
kernel sources run as the host emulation here; on the GPU box the same code path runs the CUDA
library) and one all-gather per output array reassembles the batch; the result must be bit-identical
to the single-process run (no reductions anywhere)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np, torch, torch.distributed as dist
from world_b200.api import World
from world_b200.shard import shard_ranges, all_gather_rows
from synth import synth_batch
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
fs, n = 16000, 6400
lens = [6400, 5000, 6400, 4200, 6000]
x = synth_batch(range(1, 6), fs, n).numpy()
w = World(lib_path=os.path.join({root!r}, "tests", "emu", "libworld_b200_emu.so"), array_module="numpy")
frames = [w.frames(fs, l) for l in lens]
ranges = shard_ranges(frames, world)
b, e = ranges[rank]
t, f0, fl = w.dio(np.ascontiguousarray(x[b:e]), fs, x_lengths=lens[b:e]) if e > b else (np.zeros((0, max(frames))),) * 2 + ([],)
L = max(frames)
def pad(a):
    out = np.zeros((e - b, L) + a.shape[2:]); out[:, :a.shape[1]] = a; return out
if e > b:
    f0 = w.stonemask(np.ascontiguousarray(x[b:e]), fs, t, f0, x_lengths=lens[b:e], f0_lengths=fl)
    opt = w.cheaptrick_option(fs)
    sp = w.cheaptrick(np.ascontiguousarray(x[b:e]), fs, t, f0, opt, x_lengths=lens[b:e], f0_lengths=fl)
    w.synchronize()
    for u in range(e - b):
        sp[u, fl[u]:] = 0.0
    t, f0, sp = pad(t), pad(f0), pad(sp)
else:
    sp = np.zeros((0, L, 513))
counts = [r[1] - r[0] for r in ranges]
g_f0 = all_gather_rows(dist, torch.from_numpy(f0), counts).numpy()
g_t = all_gather_rows(dist, torch.from_numpy(t), counts).numpy()
g_sp = all_gather_rows(dist, torch.from_numpy(sp), counts).numpy()
if rank == 0:
    np.savez({out!r}, f0=g_f0, t=g_t, sp=g_sp, ranges=np.array(ranges))
dist.destroy_process_group()
'''


def test_shard_ranges_balance():
    sys.path.insert(0, ROOT)
    from world_b200.shard import shard_ranges
    assert shard_ranges([10] * 8, 4) == [(0, 2), (2, 4), (4, 6), (6, 8)]
    r = shard_ranges([100, 1, 1, 1, 100], 2)
    assert r[0][0] == 0 and r[-1][1] == 5 and all(a[1] == b[0] for a, b in zip(r, r[1:]))
    assert shard_ranges([5], 3)[-1][1] == 1


def test_two_rank_gather_equals_single_process(emu, tmp_path):
    out = str(tmp_path / "gathered.npz")
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, out=out))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                           "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)], env=env,
                          timeout=600)
    got = np.load(out)
    from synth import synth_batch
    fs, n = 16000, 6400
    lens = [6400, 5000, 6400, 4200, 6000]
    x = synth_batch(range(1, 6), fs, n).numpy()
    t, f0, fl = emu.dio(x, fs, x_lengths=lens)
    f0 = emu.stonemask(x, fs, t, f0, x_lengths=lens, f0_lengths=fl)
    sp = emu.cheaptrick(x, fs, t, f0, emu.cheaptrick_option(fs), x_lengths=lens, f0_lengths=fl)
    emu.synchronize()
    for u in range(5):
        sp[u, fl[u]:] = 0.0
    assert np.array_equal(got["t"], t)
    assert np.array_equal(got["f0"], f0)
    assert np.array_equal(got["sp"], sp)
