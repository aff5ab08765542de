#!/usr/bin/env python
"""bench.py -- GTEPS of the hot path (send -> SpMSpV -> apply) on synthetic RMAT.

Workload (BASELINE.json metric / configs[2], north_star target): PageRank on RMAT scale-26
(a,b,c = .57,.19,.19, edge factor 16, seed 1, duplicates kept), generated on the device.
One "step" = run_graph_program(PageRank, ITERS iterations) over the resident graph.

  value    whole-job GTEPS = nnz * ITERS * steps / device time, inputs resident in HBM
  e2e      the same through the C ABI with HOST buffers: vertex properties uploaded from pinned
           memory (setVertexproperty for all), run, results downloaded (getVertexproperty for all)
  roofline dominant kernels = the SpMSpV pass (k_heavy + k_sell): algorithmic bytes
           nnz*(sizeof(E)+4) + |active|*sizeof(M) per pass  /  CUDA-event time of the pass
  cpu_baseline  the UNMODIFIED reference (oracle/_ref) on this box's host cores, bounded sample

`--impl reference` times only the reference CPU path (rank 0) and prints the same JSON shape.
N > 1: one process per GPU (torchrun); rows are sharded one tile-row per rank, the message vector
is all-gathered over NCCL every iteration.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="graphmat_b200", choices=["graphmat_b200", "reference"])
    ap.add_argument("--scale", type=int, default=int(os.environ.get("GM_BENCH_SCALE", "26")))
    ap.add_argument("--iters", type=int, default=10, help="PageRank iterations per step")
    ap.add_argument("--threads", type=int, default=4, help="ref_threads of the logical layout")
    ap.add_argument("--cpu-scale", type=int, default=int(os.environ.get("GM_BENCH_CPU_SCALE", "22")))
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-bfs", action="store_true")
    ap.add_argument("--heavy", type=int, default=0)
    return ap.parse_args()


def measured_traffic(scale, world):
    """DRAM bytes of one SpMSpV pass from the committed ncu --set full capture (profiles/), same workload only."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic_pagerank_rmat26.json")))
        if scale == 26 and world == 1:
            return float(t["dram_bytes_per_pass"]), t["source"]
    except Exception:
        pass
    return None, None


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6:
                    self.rows.append(f)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None, "reasons": reasons,
                "samples": len(self.rows)}


def cpu_reference(scale, iters, steps, warmup):
    """The unmodified reference on the host cores: PageRank, same generator, bounded scale."""
    from graphmat_b200 import capi
    from oracle import ref
    cores = os.cpu_count() or 1
    n, s, d, _ = capi.rmat_edges(scale, 16, seed=1)
    sess = ref.PageRankSession(n, s, d, None, threads=cores)
    times = []
    for k in range(warmup + steps):
        it, ms = sess.run(iters)
        if k >= warmup:
            times.append(ms)
    sess.close()
    nnz = len(s)
    total_ms = sum(times)
    gteps = nnz * iters * len(times) / (total_ms * 1e-3) / 1e9
    return gteps, total_ms / len(times), cores, nnz


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    gteps, ms, cores, nnz = cpu_reference(args.cpu_scale, args.iters, args.steps, min(args.warmup, 1))
    sample = "PageRank RMAT scale-%d (%d edges), %d iterations per step, 1 rank (stub MPI) x %d OpenMP threads" % (
        args.cpu_scale, nnz, args.iters, cores)
    line = {"impl": "reference", "metric": "GTEPS (PageRank, nnz*iterations/time of run_graph_program)", "value": gteps,
            "unit": "GTEPS", "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "PageRank on synthetic RMAT scale-%d" % args.scale, "sample": sample},
            "cpu_baseline": {"value": gteps, "unit": "GTEPS", "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": gteps, "unit": "GTEPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    from graphmat_b200 import capi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    capi._check(capi.lib().gm_set_device(C.c_int(local)), "gm_set_device")
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    t_build = time.time()
    G = capi.Graph.rmat(args.scale, capi.PR_DTYPE, seed=1, threads=args.threads, rank=rank, world=world,
                        heavy_threshold=args.heavy)
    gv = G.view()
    n, nnz = gv.nvertices, gv.nnz
    tmp = capi.Vectors(G, capi.PROG_PAGERANK)
    if world > 1:
        from graphmat_b200 import exchange
        exchange.attach(G, tmp, dist)
    build_s = time.time() - t_build

    init = np.zeros(1, capi.PR_DTYPE)
    init["pagerank"], init["degree"] = 0.3, 0
    G.set_all_vertexproperty(init[0])
    G.set_all_active()
    G.run(capi.PROG_DEGREE, None, 1)  # out-degrees (src/PageRank.cpp:133-139), outside the timed region
    state = capi.PageRankState(0.3)

    def step():
        G.set_all_active()
        return G.run(capi.PROG_PAGERANK, state, args.iters, tmp)

    for _ in range(max(args.warmup, 3)):
        step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    dev_ms = spmv_ms = 0.0
    launches = 0
    for _ in range(args.steps):
        st = step()
        dev_ms += st.ms_total
        spmv_ms += st.ms_spmv
        launches += st.kernel_launches
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    sampler.stop_flag = True
    # max over ranks of the device time
    if dist is not None:
        t = torch.tensor([dev_ms, spmv_ms, wall_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, spmv_ms, wall_ms = t.tolist()
    passes = args.iters * args.steps
    gteps = nnz * passes / (dev_ms * 1e-3) / 1e9

    # ---- e2e: host buffers in pinned memory, copies inside the timed region ----
    vdt = capi.PR_DTYPE
    host_in = torch.empty(n * vdt.itemsize, dtype=torch.uint8).pin_memory()
    host_out = torch.empty(n * vdt.itemsize, dtype=torch.uint8).pin_memory()
    vp0 = G.get_vertexproperties()
    vp0["pagerank"] = 0.3
    np.frombuffer(host_in.numpy(), dtype=vdt)[:] = vp0
    e2e_steps = max(1, min(args.steps, 3))

    def e2e_step():
        G.set_vertexproperties_ptr(host_in.data_ptr())
        G.set_all_active()
        G.run(capi.PROG_PAGERANK, state, args.iters, tmp)
        G.get_vertexproperties_ptr(host_out.data_ptr())

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    if dist is not None:
        t = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = t.item()
    e2e_gteps = nnz * args.iters * e2e_steps / (e2e_ms * 1e-3) / 1e9

    # ---- BFS on RMAT-22 (configs[1]), reported beside the headline ----
    bfs = None
    if not args.no_bfs and world == 1:
        from graphmat_b200 import apps
        Gb = capi.Graph.rmat(22, capi.BFS_DTYPE, seed=1, threads=args.threads, build_mask=2)
        Gb.push_ready(1)  # column-major companion of the sparse-frontier path: graph construction, not BFS time
        src0 = Gb.first_source()
        nb = Gb.nvertices
        vp = np.zeros(nb, capi.BFS_DTYPE)
        best = None
        for _ in range(4):
            vp["depth"] = 0xFFFFFFFF
            vp["parent"] = np.uint64(0xFFFFFFFFFFFFFFFF)
            vp["id"] = np.arange(1, nb + 1, dtype=np.uint64)
            vp["depth"][src0 - 1] = 0
            Gb.set_vertexproperties(vp)
            Gb.set_all_inactive()
            Gb.set_active(src0)
            stb = Gb.run(capi.PROG_BFS, capi.BFSState(1), capi.UNTIL_CONVERGENCE)
            best = stb.ms_total if best is None else min(best, stb.ms_total)
        reach = int(Gb.reduce(capi.REDUCE_REACHABLE))
        bfs = {"workload": "BFS RMAT scale-22", "gteps": Gb.nnz / (best * 1e-3) / 1e9, "ms": best,
               "iterations": stb.iterations, "reachable": reach, "source": src0,
               "push_passes": int(stb.push_passes), "entries_swept": int(stb.edges_processed)}
        Gb.close()

    if rank == 0:
        peak, peak_src = peaks()
        # algorithmic bytes per SpMSpV pass: nnz*(sizeof(E)+sizeof(idx)) + |active|*sizeof(M)
        alg_bytes = nnz * (4 + 4) + n * 4
        ms_per_pass = spmv_ms / passes
        traffic, traffic_src = measured_traffic(args.scale, world)
        achieved = alg_bytes / world / (ms_per_pass * 1e-3) / 1e9  # per GPU
        line = {
            "metric": "GTEPS (PageRank, nnz*iterations/time of run_graph_program)", "value": gteps, "unit": "GTEPS",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "PageRank on synthetic RMAT scale-%d" % args.scale, "vertices": n, "edges": nnz,
                       "iterations_per_step": args.iters, "edge_factor": 16, "rmat": "a,b,c=.57,.19,.19 seed=1",
                       "ref_threads": args.threads, "sharding": "one tile-row per GPU, x all-gather" if world > 1 else "1 GPU",
                       "l2": "inputs (%.1f GB index stream) larger than L2, no flush" % (nnz * 4 / 1e9),
                       "build_seconds": round(build_s, 2)},
            "wall_ms_per_step": wall_ms / args.steps,
            "e2e": {"value": e2e_gteps, "unit": "GTEPS", "h2d_bytes_per_step": n * vdt.itemsize,
                    "d2h_bytes_per_step": n * vdt.itemsize, "ms_per_step": e2e_ms / e2e_steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "kernel": "SpMSpV pass (k_heavy + k_sell)", "ms_per_launch": ms_per_pass,
                         "algorithmic_bytes": alg_bytes // world, "peak_source": peak_src,
                         "spmv_share_of_step": spmv_ms / dev_ms},
            "clocks": sampler.summary(),
        }
        if bfs:
            line["bfs"] = bfs
        if not args.no_cpu and world == 1:
            try:
                cg, cms, cores, cnnz = cpu_reference(args.cpu_scale, args.iters, 2, 1)
                line["cpu_baseline"] = {"value": cg, "unit": "GTEPS", "cores": cores, "kind": "reference",
                                        "sample": "PageRank RMAT scale-%d (%d edges), %d iterations, unmodified reference, "
                                                  "1 rank (stub MPI) x %d OpenMP threads" % (args.cpu_scale, cnnz, args.iters, cores)}
            except Exception as e:  # the reference .so is test infrastructure; report, do not hide
                line["cpu_baseline"] = {"value": None, "unit": "GTEPS", "cores": os.cpu_count(), "kind": "reference",
                                        "sample": "unavailable: %s" % e}
        print(json.dumps(line))
    tmp.close()
    G.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
