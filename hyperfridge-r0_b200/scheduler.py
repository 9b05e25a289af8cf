"""Host-side partitioning of independent segments over ranks / GPUs (no data-path collective).

Used by bench.py under torchrun (one rank per GPU) and mirrored by the in-process pool (hfb200_pool).  Segments of a
session are independent: rank r proves jobs r, r + world, r + 2*world, ... of the list sorted
longest-first, and every job carries its own blinding seed = f(global seed, job id, segment index), so the seals do
not depend on how many GPUs took part."""


def job_seed(global_seed, job_id, segment_index):
    """Deterministic per-segment blinding seed (SplitMix64 of the triple)."""
    z = (global_seed * 0x9E3779B97F4A7C15 + job_id * 0xBF58476D1CE4E5B9 + segment_index * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return z ^ (z >> 31)


def make_batch(n_statements=64, base_segments=30, spread=16, po2=20):
    """BASELINE.json configs[3]: 64 synthetic camt53 statements, statement i has 30 + (i mod 16) segments of po2."""
    jobs = []
    for i in range(n_statements):
        for s in range(base_segments + (i % spread)):
            jobs.append({"statement": i, "segment": s, "po2": po2})
    return jobs


def shard(jobs, rank, world):
    """Round-robin over the longest-first order: balanced to within one job, deterministic."""
    order = sorted(range(len(jobs)), key=lambda i: (-jobs[i]["po2"], i))
    return [jobs[i] for i in order[rank::world]]
