"""Multi-GPU host logic on CPU: segment sharding is deterministic and complete, seals do not depend on the number
of workers, and the N>1 bench plumbing (one process per GPU, no data-path collective) works under a world_size-2
gloo group.  Kernels run through the host emulator here; the same paths run on real GPUs in test_gpu_parity."""
import os
import sys
import numpy as np
import pytest
from conftest import SMALL, make_segment

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sched():
    import importlib
    import hfb200_loader
    hfb200_loader.load()
    return importlib.import_module("hyperfridge_r0_b200.scheduler")


def test_shard_partition_is_complete_and_balanced():
    s = _sched()
    jobs = s.make_batch(64, 30, 16, 20)
    assert len(jobs) == sum(30 + (i % 16) for i in range(64))
    for world in (1, 2, 4, 8):
        parts = [s.shard(jobs, r, world) for r in range(world)]
        flat = [(j["statement"], j["segment"]) for p in parts for j in p]
        assert sorted(flat) == sorted((j["statement"], j["segment"]) for j in jobs)
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    # seeds depend on (global seed, job, segment) only
    assert s.job_seed(1, 2, 3) == s.job_seed(1, 2, 3) != s.job_seed(1, 2, 4)
    mixed = [{"statement": 0, "segment": i, "po2": p} for i, p in enumerate([16, 20, 18, 20])]
    assert [j["po2"] for j in s.shard(mixed, 0, 1)] == [20, 20, 18, 16]  # longest first


def test_pool_seals_independent_of_worker_count(pkg, emu_lib, orc):
    jobs, expect = [], []
    for i, po2 in enumerate([12, 13, 12, 12, 13]):
        cir, g, code, data = make_segment(orc, SMALL, po2, trace_seed=100 + i)
        jobs.append((po2, g, code, data, 5 + i))
        expect.append(cir.prove(po2, g, code, data, 5 + i)[0])
    for devices, per in (((0,), 1), ((0, 0), 1), ((0, 0, 0), 2)):
        with pkg.Pool(devices=devices, contexts_per_device=per, max_po2=13, circuit=SMALL, lib=emu_lib) as pool:
            seals, devs, ms = pool.prove(jobs, 40000)
            assert all(len(a) == len(b) and (a == b).all() for a, b in zip(seals, expect))
    with pkg.Pool(devices=(0,), max_po2=12, circuit=SMALL, lib=emu_lib) as pool:
        with pytest.raises(pkg.Hfb200Error):  # po2 13 > max_po2: the job's error is reported, nothing is silently skipped
            pool.prove(jobs, 40000)


def _rank_main(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import importlib
    import hfb200_loader
    import oracle
    pkg = hfb200_loader.load()
    sched = importlib.import_module("hyperfridge_r0_b200.scheduler")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = pkg.load_library(os.path.join(ROOT, "tests", "emu", "libhfb200_emu.so"))
    jobs = [{"statement": i // 2, "segment": i % 2, "po2": 12} for i in range(6)]
    mine = sched.shard(jobs, rank, world)
    cir = oracle.Circuit(*SMALL)
    out = []
    with pkg.Context(0, 12, SMALL, lib=lib) as ctx:
        for j in mine:
            seed = sched.job_seed(7, j["statement"], j["segment"])
            g = ctx.witgen_synth(12, 1000 + 10 * j["statement"] + j["segment"], seed)
            seal = ctx.prove_resident(seed)
            out.append((j["statement"], j["segment"], int(oracle.hash_elems(seal % oracle.P)[0])))
    # the only communication: timing barrier / max-over-ranks, and gathering results on rank 0
    t = torch.tensor([float(len(mine))])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    if rank == 0:
        q.put(sorted(x for part in gathered for x in part))
    dist.destroy_process_group()


def test_two_rank_gloo_run_matches_single_process(pkg, emu_lib, orc):
    import importlib
    import torch.multiprocessing as mp
    sched = importlib.import_module("hyperfridge_r0_b200.scheduler")
    ctx_mp = mp.get_context("spawn")
    q = ctx_mp.Queue()
    port = 29500 + (os.getpid() % 1000)
    procs = [ctx_mp.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    got = q.get(timeout=240)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    # single-process reference over the same job list
    expect = []
    with pkg.Context(0, 12, SMALL, lib=emu_lib) as ctx:
        for i in range(6):
            st, sg = i // 2, i % 2
            seed = sched.job_seed(7, st, sg)
            ctx.witgen_synth(12, 1000 + 10 * st + sg, seed)
            expect.append((st, sg, int(orc.hash_elems(ctx.prove_resident(seed) % orc.P)[0])))
    assert got == sorted(expect)


def test_reference_arm_contract(orc):
    """`bench.py --impl reference`: one JSON line with the contract's keys on rank 0, nothing on the other ranks."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--sample-po2", "13"]
    env = dict(os.environ, RANK="0", WORLD_SIZE="2")
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "segments/s" and d["higher_is_better"] is True and d["n_gpus"] == 2
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "segments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and "po2=20" in d["metric"]
    env["RANK"] = "1"
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == ""
