"""bench_workloads.py -- the other BASELINE.json configurations behind `bench.py --workload ...`:

  bfs            BFS on RMAT scale-22 (config 2), 1..N GPUs
  sssp           SSSP on weighted RMAT scale-24 (config 5's graph), weights uniform 1..127
  deltastepping  DeltaStepping (src/DeltaStepping.cpp:124-198) on the same graph, delta = 16
  sgd            SGD collaborative filtering, 10 M users x 1 M items, K = 32 (config 4); --ratings sets nnz

Same JSON contract as bench.py's PageRank line: `value` = whole-job GTEPS on the device clock (max over ranks),
`e2e` = the same with the vertex properties uploaded from / downloaded to pinned host memory inside the timed
region, `roofline` by SURVEY.md 8(d)'s formulas (SGD: formula and gather-inclusive), and `parity`: the same
program on the same N GPUs at a reduced scale against the CPU oracle (bit-exact; SGD within 1e-6).
"""
import json
import os
import time

import numpy as np


def _small_parity(c, args, B, what):
    """reduced-scale run on the SAME ranks / exchange against the oracle (rank 0 compares)"""
    from graphmat_b200 import apps, capi, exchange
    from oracle import port
    rank, world, dist = c.rank, c.world, c.dist
    threads = 4
    res = {"checked": None, "ok": None}
    if what in ("bfs", "sssp", "deltastepping"):
        scale = 14
        n, s, d, w = capi.rmat_edges(scale, 16, seed=1, weight_max=127)
        src0 = int(s.min())
    if what == "bfs":
        G = capi.Graph.from_edges(n, s, d, None, capi.BFS_DTYPE, threads=threads, rank=rank, world=world)
        B.attach_exchange(c, G, args)
        tmp = capi.Vectors(G, capi.PROG_BFS)
        vp = np.zeros(n, capi.BFS_DTYPE)
        vp["depth"] = 0xFFFFFFFF
        vp["parent"] = np.uint64(0xFFFFFFFFFFFFFFFF)
        vp["id"] = np.arange(1, n + 1, dtype=np.uint64)
        vp["depth"][src0 - 1] = 0
        exchange.upload_vp(G, vp, rank)
        G.set_all_inactive()
        G.set_active(src0)
        st = G.run(capi.PROG_BFS, capi.BFSState(1), capi.UNTIL_CONVERGENCE, tmp)
        got = exchange.collect_vp(G, capi.BFS_DTYPE, n, dist, rank, world)
        od, op, oit, _ = port.bfs(n, s, d, src0, threads=threads)
        res = {"checked": "BFS RMAT-%d depth+parent vs oracle on %d GPU(s)" % (scale, world),
               "ok": bool((got["depth"] == od).all() and (got["parent"] == op).all() and st.iterations == oit)}
        tmp.close(); G.close()
    elif what == "sssp":
        G = capi.Graph.from_edges(n, s, d, w, capi.SSSP_DTYPE, threads=threads, rank=rank, world=world)
        B.attach_exchange(c, G, args)
        tmp = capi.Vectors(G, capi.PROG_SSSP)
        _sssp_init(capi, G, src0)
        st = G.run(capi.PROG_SSSP, None, capi.UNTIL_CONVERGENCE, tmp)
        got = exchange.collect_vp(G, capi.SSSP_DTYPE, n, dist, rank, world)
        odist, osit, _ = port.sssp(n, s, d, w, src0, threads=threads)
        res = {"checked": "SSSP RMAT-%d distances vs oracle on %d GPU(s)" % (scale, world),
               "ok": bool((got["distance"] == odist).all() and st.iterations == osit)}
        tmp.close(); G.close()
    elif what == "deltastepping":
        G, G2, tmp = _ds_graphs(c, args, B, capi, n, s, d, w, 16, threads)
        nb, _ = _ds_run(c, capi, G, G2, tmp, 16, src0)
        got = exchange.collect_vp(G, capi.DS_DTYPE, n, dist, rank, world)
        odd, odb, onb, _ = port.deltastepping(n, s, d, w, 16, src0, threads=threads)
        res = {"checked": "DeltaStepping RMAT-%d distance+bucket vs oracle on %d GPU(s)" % (scale, world),
               "ok": bool((got["distance"] == odd).all() and (got["bucket"] == odb).all() and nb == onb)}
        tmp.close(); G2.close(); G.close()
    elif what == "sgd":
        m_users, n_items, K, nr = 3000, 400, 32, 60000
        u, it_, r_ = B.synth_ratings(m_users, n_items, nr)
        nv = m_users + n_items
        dt = capi.latent_dtype(K)
        p_sgd, _ = capi.SGD_PROGRAMS[K]
        G = capi.Graph.from_edges(nv, u, it_, r_, dt, threads=threads, rank=rank, world=world)
        B.attach_exchange(c, G, args)
        tmp = capi.Vectors(G, p_sgd)
        vp = np.zeros(nv, dt)
        vp["lv"] = apps.sgd_init(nv, K)
        exchange.upload_vp(G, vp, rank)
        G.set_all_active()
        G.run(p_sgd, capi.SGDState(0.001, 0.00000035), 10, tmp)
        got = exchange.collect_vp(G, dt, nv, dist, rank, world)
        olv, _, _ = port.sgd(m_users, nv, u, it_, r_, K=K, iterations=10, threads=threads)
        err = float((np.abs(got["lv"] - olv) / np.maximum(np.abs(olv), 1e-300)).max())
        res = {"checked": "SGD K=32, %d ratings, 10 iterations vs oracle on %d GPU(s)" % (nr, world), "max_rel_err": err,
               "tolerance": 1e-6, "ok": err <= 1e-6}
        tmp.close(); G.close()
    return res


def _sssp_init(capi, G, src0):
    inf = np.zeros(1, capi.SSSP_DTYPE)
    inf["distance"] = 0xFFFFFFFF
    G.set_all_vertexproperty(inf[0])
    G.set_all_inactive()
    G.set_vertexproperty(src0, np.zeros(1, capi.SSSP_DTYPE)[0])
    G.set_active(src0)


def _ds_graphs(c, args, B, capi, n, s, d, w, delta, threads):
    light = w <= delta  # filter_edges(less_than_delta), src/DeltaStepping.cpp:136-137
    G = capi.Graph.from_edges(n, s[light], d[light], w[light], capi.DS_DTYPE, threads=threads, rank=c.rank, world=c.world,
                              build_mask=2)
    G2 = capi.Graph.from_edges(n, s[~light], d[~light], w[~light], capi.DS_DTYPE, threads=threads, rank=c.rank,
                               world=c.world, order_like=G, build_mask=2)
    G2.share_vertexproperty(G)
    B.attach_exchange(c, G, args)
    B.attach_exchange(c, G2, args)
    tmp = capi.Vectors(G, capi.PROG_DELTASTEPPING)
    return G, G2, tmp


def _ds_run(c, capi, G, G2, tmp, delta, src0):
    """the bucket loop of src/DeltaStepping.cpp:166-177 -> (buckets, device ms inside run_graph_program)"""
    init = np.zeros(1, capi.DS_DTYPE)
    init["distance"], init["bucket"] = 0xFFFFFFFF, 0x7FFFFFFF
    G.set_all_vertexproperty(init[0])
    G.set_all_inactive()
    G.set_vertexproperty(src0, np.zeros(1, capi.DS_DTYPE)[0])
    G.set_active(src0)
    state = capi.DeltaSteppingState(delta, 0)
    dev_ms = 0.0
    while True:
        G.set_all_active()
        dev_ms += G.run(capi.PROG_DELTASTEPPING, state, capi.UNTIL_CONVERGENCE, tmp).ms_total
        G2.set_all_active()
        dev_ms += G2.run(capi.PROG_DELTASTEPPING, state, 1, tmp).ms_total
        state.bid += 1
        left = G.reduce(capi.REDUCE_BUCKET_NOT_EMPTY, state.bid)
        if c.dist is not None:
            t = c.torch.tensor([left], dtype=c.torch.float64, device="cuda")
            c.dist.all_reduce(t)
            left = t.item()
        if left == 0:
            break
    return state.bid, dev_ms


def run(c, args, B):
    capi, torch = c.capi, c.torch
    rank, world, dist = c.rank, c.world, c.dist
    what = args.workload
    parity = _small_parity(c, args, B, what)
    peak, peak_src = B.peaks()
    sampler = B.ClockSampler(c.local)
    extra = {}
    t_build = time.time()

    if what in ("bfs", "sssp"):
        dt = capi.BFS_DTYPE if what == "bfs" else capi.SSSP_DTYPE
        G = capi.Graph.rmat(args.scale, dt, seed=1, weight_max=0 if what == "bfs" else 127, weight_seed=2, threads=args.threads,
                            rank=rank, world=world, build_mask=2)
        sharding = B.attach_exchange(c, G, args)
        prog = capi.PROG_BFS if what == "bfs" else capi.PROG_SSSP
        tmp = capi.Vectors(G, prog)
        G.push_ready(1)
        n, nnz = G.nvertices, G.nnz
        src0 = G.first_source()
        build_s = time.time() - t_build
        sliced = world > 1 and G.peers_enabled()
        lo, hi = G.slice_range(rank) if sliced else (0, n)
        if what == "bfs":
            vp = np.zeros(hi - lo, dt)
            vp["depth"] = 0xFFFFFFFF
            vp["parent"] = np.uint64(0xFFFFFFFFFFFFFFFF)
            vp["id"] = np.arange(lo + 1, hi + 1, dtype=np.uint64)
            if lo < src0 <= hi:
                vp["depth"][src0 - 1 - lo] = 0
        else:
            vp = np.zeros(hi - lo, dt)
            vp["distance"] = 0xFFFFFFFF
            if lo < src0 <= hi:
                vp["distance"][src0 - 1 - lo] = 0
        host_in = torch.empty(max(1, vp.nbytes), dtype=torch.uint8).pin_memory()
        host_out = torch.empty(max(1, vp.nbytes), dtype=torch.uint8).pin_memory()
        np.frombuffer(host_in.numpy(), dtype=dt)[:len(vp)] = vp

        def reset():
            if sliced:
                G.set_vertexproperties_slice_ptr(host_in.data_ptr())
            else:
                G.set_vertexproperties_ptr(host_in.data_ptr())
            G.set_all_inactive()
            G.set_active(src0)

        def run_once():
            return G.run(prog, capi.BFSState(1) if what == "bfs" else None, capi.UNTIL_CONVERGENCE, tmp)

        for _ in range(max(args.warmup, 3)):
            reset()
            run_once()
        if rank == 0:
            sampler.start()
        dev_ms = spmv_ms = 0.0
        launches = 0
        B.barrier(c)
        for _ in range(args.steps):
            reset()
            st = run_once()
            dev_ms += st.ms_total
            spmv_ms += st.ms_spmv
            launches += st.kernel_launches
        B.barrier(c)
        sampler.stop_flag = True
        dev_ms, spmv_ms = B.max_over_ranks(c, [dev_ms, spmv_ms])
        # e2e: upload + run + download inside the clock
        B.barrier(c)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            reset()
            run_once()
            if sliced:
                G.get_vertexproperties_slice_ptr(host_out.data_ptr())
            else:
                G.get_vertexproperties_ptr(host_out.data_ptr())
        B.barrier(c)
        e2e_ms, = B.max_over_ranks(c, [(time.perf_counter() - t0) * 1e3])
        reach = G.reduce(capi.REDUCE_REACHABLE)
        if dist is not None:
            t = torch.tensor([reach], dtype=torch.float64, device="cuda")
            dist.all_reduce(t)
            reach = t.item()
        gteps = nnz * args.steps / (dev_ms * 1e-3) / 1e9
        e2e = nnz * args.steps / (e2e_ms * 1e-3) / 1e9
        swept = int(st.edges_processed)
        alg = nnz * 8 + int(reach) * (8 if what == "bfs" else 4)   # one full sweep of the matrix + one message per reached vertex
        moved = n * dt.itemsize if (sliced or world == 1) else n * dt.itemsize * world
        roof = {"bound": "hbm", "unit": "GB/s", "peak": peak, "peak_source": peak_src,
                "algorithmic_bytes": alg // world, "ms_per_launch": dev_ms / args.steps,
                "achieved": alg / world / (dev_ms / args.steps * 1e-3) / 1e9,
                "kernel": "whole run_graph_program (%d iterations, %d sparse-frontier passes)" % (st.iterations, st.push_passes),
                "traffic": None,
                "note": "SURVEY 8(d): nnz*(sizeof(E)+4) + |reached|*sizeof(M) for the whole traversal; the run is %d launch-"
                        "latency-bound iterations, so the fraction says how far a traversal is from one streaming sweep" % st.iterations}
        roof["frac"] = roof["achieved"] / peak
        extra = {"iterations": st.iterations, "push_passes": int(st.push_passes), "entries_swept": swept, "reachable": int(reach),
                 "source": src0}
        tmp.close(); G.close()
        dtype = "u64" if what == "bfs" else "u32"

    elif what == "deltastepping":
        n, s, d, w = capi.rmat_edges(args.scale, 16, seed=1, weight_max=127, weight_seed=2)
        src0 = int(s.min())
        nnz = len(s)
        delta = 16
        G, G2, tmp = _ds_graphs(c, args, B, capi, n, s, d, w, delta, args.threads)
        G.push_ready(1)
        G2.push_ready(1)
        sharding = "one tile-row per GPU" if world > 1 else "1 GPU"
        if world > 1:
            sharding = "peer memory" if G.peers_enabled() else "NCCL callbacks"
        build_s = time.time() - t_build
        del s, d, w
        for _ in range(max(args.warmup, 3)):
            nb, _ = _ds_run(c, capi, G, G2, tmp, delta, src0)
        if rank == 0:
            sampler.start()
        B.barrier(c)
        t0 = time.perf_counter()
        dev_ms = 0.0
        for _ in range(args.steps):
            nb, ms = _ds_run(c, capi, G, G2, tmp, delta, src0)
            dev_ms += ms
        B.barrier(c)
        wall_ms = (time.perf_counter() - t0) * 1e3
        sampler.stop_flag = True
        dev_ms, wall_ms = B.max_over_ranks(c, [dev_ms, wall_ms])
        host_out = torch.empty(max(1, n * capi.DS_DTYPE.itemsize), dtype=torch.uint8).pin_memory()
        sliced = world > 1 and G.peers_enabled()
        B.barrier(c)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            _ds_run(c, capi, G, G2, tmp, delta, src0)
            if sliced:
                G.get_vertexproperties_slice_ptr(host_out.data_ptr())
            else:
                G.get_vertexproperties_ptr(host_out.data_ptr())
        B.barrier(c)
        e2e_ms, = B.max_over_ranks(c, [(time.perf_counter() - t0) * 1e3])
        # the reference times the whole bucket loop on the wall clock (src/DeltaStepping.cpp:165-180): so does `value`
        gteps = nnz * args.steps / (wall_ms * 1e-3) / 1e9
        e2e = nnz * args.steps / (e2e_ms * 1e-3) / 1e9
        moved = n * capi.DS_DTYPE.itemsize if (sliced or world == 1) else n * capi.DS_DTYPE.itemsize * world
        alg = nnz * 8 + n * 4
        roof = {"bound": "hbm", "unit": "GB/s", "peak": peak, "peak_source": peak_src, "algorithmic_bytes": alg // world,
                "ms_per_launch": wall_ms / args.steps, "achieved": alg / world / (wall_ms / args.steps * 1e-3) / 1e9,
                "kernel": "whole bucket loop (%d buckets)" % nb, "traffic": None,
                "note": "one sweep of both edge sets + one message per vertex; device time inside run_graph_program %.2f ms of "
                        "%.2f ms wall per loop" % (dev_ms / args.steps, wall_ms / args.steps)}
        roof["frac"] = roof["achieved"] / peak
        extra = {"buckets": int(nb), "delta": delta, "source": src0, "device_ms_per_step": dev_ms / args.steps}
        dev_ms = wall_ms
        launches = 0
        tmp.close(); G2.close(); G.close()
        dtype = "u32"

    else:  # sgd
        K = 32
        m_users, n_items = 10_000_000, 1_000_000
        nr = args.ratings
        nv = m_users + n_items
        gen = torch.Generator(device="cuda")
        gen.manual_seed(3)
        # user uniform, item Zipf(1.0), rating uniform 1..5 -- generated on the device, same on every rank
        u = torch.randint(1, m_users + 1, (nr,), generator=gen, device="cuda", dtype=torch.int32)
        wz = 1.0 / torch.arange(1, n_items + 1, device="cuda", dtype=torch.float64)
        cdf = torch.cumsum(wz / wz.sum(), 0)
        it = (torch.searchsorted(cdf, torch.rand(nr, generator=gen, device="cuda", dtype=torch.float64)).clamp_(max=n_items - 1)
              + 1 + m_users).to(torch.int32)
        r = torch.randint(1, 6, (nr,), generator=gen, device="cuda", dtype=torch.int32)
        del cdf, wz
        dt = capi.latent_dtype(K)
        p_sgd, _ = capi.SGD_PROGRAMS[K]
        G = capi.Graph.from_device_edges(nv, nr, u.data_ptr(), it.data_ptr(), r.data_ptr(), dt, threads=args.threads, rank=rank,
                                         world=world)
        del u, it, r
        torch.cuda.empty_cache()
        sharding = B.attach_exchange(c, G, args)
        tmp = capi.Vectors(G, p_sgd)
        n, nnz = nv, nr
        build_s = time.time() - t_build
        sliced = world > 1 and G.peers_enabled()
        lo, hi = G.slice_range(rank) if sliced else (0, nv)
        host_in = torch.empty(max(1, (hi - lo) * dt.itemsize), dtype=torch.uint8).pin_memory()
        host_out = torch.empty(max(1, (hi - lo) * dt.itemsize), dtype=torch.uint8).pin_memory()
        vp = np.frombuffer(host_in.numpy(), dtype=dt)[:hi - lo]
        vp["lv"] = np.random.default_rng(5 + rank).random((hi - lo, K))
        vp["sqerr"] = 0.0
        state = capi.SGDState(0.001, 0.00000035)

        def upload():
            if sliced:
                G.set_vertexproperties_slice_ptr(host_in.data_ptr())
            else:
                G.set_vertexproperties_ptr(host_in.data_ptr())

        upload()

        def step():
            G.set_all_active()
            return G.run(p_sgd, state, args.iters, tmp)

        dev_ms, spmv_ms, wall_ms, launches, sgd_clocks = B.timed_steps(c, args, step)
        B.barrier(c)
        t0 = time.perf_counter()
        esteps = max(1, min(args.steps, 2))
        for _ in range(esteps):
            upload()
            step()
            if sliced:
                G.get_vertexproperties_slice_ptr(host_out.data_ptr())
            else:
                G.get_vertexproperties_ptr(host_out.data_ptr())
        B.barrier(c)
        e2e_ms, = B.max_over_ranks(c, [(time.perf_counter() - t0) * 1e3])
        passes = 2 * args.iters * args.steps
        gteps = nnz * passes / (dev_ms * 1e-3) / 1e9
        e2e = nnz * 2 * args.iters * esteps / (e2e_ms * 1e-3) / 1e9
        e2e_ms = e2e_ms / esteps * args.steps
        moved = nv * dt.itemsize if (sliced or world == 1) else nv * dt.itemsize * world
        alg = nnz * 8 + nv * dt.itemsize             # per pass, SURVEY 8(d) formula
        alg_gather = alg + nnz * dt.itemsize         # + the destination's vertex property per edge (my_spmspv3)
        ms_pass = spmv_ms / passes
        roof = {"bound": "hbm", "unit": "GB/s", "peak": peak, "peak_source": peak_src, "algorithmic_bytes": alg // world,
                "ms_per_launch": ms_pass, "achieved": alg / world / (ms_pass * 1e-3) / 1e9,
                "achieved_gather_inclusive": alg_gather / world / (ms_pass * 1e-3) / 1e9,
                "kernel": "one SpMSpV3 pass (k_sell/k_heavy with the destination's vertex property)", "traffic": None,
                "note": "formula nnz*(sizeof(E)+4) + n*sizeof(M) undercounts 264-byte payloads; gather-inclusive adds nnz*264 B"}
        roof["frac"] = roof["achieved"] / peak
        roof["frac_gather_inclusive"] = roof["achieved_gather_inclusive"] / peak
        extra = {"ratings": nr, "K": K, "iterations_per_step": args.iters}
        sampler.stop_flag = True
        sampler.summary = lambda: sgd_clocks
        tmp.close(); G.close()
        dtype = "f64"

    if rank != 0:
        return
    line = {"metric": B.METRIC[what], "value": gteps, "unit": "GTEPS", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": dict({"workload": B.workload_name(args), "vertices": int(n), "edges": int(nnz), "ref_threads": args.threads,
                            "sharding": sharding, "build_seconds": round(build_s, 2),
                            "l2": "vectors and index streams larger than L2, no flush"}, **extra),
            "e2e": {"value": e2e, "unit": "GTEPS", "h2d_bytes_per_step": int(moved), "d2h_bytes_per_step": int(moved),
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches), "roofline": roof, "clocks": sampler.summary(), "parity": parity}
    print(json.dumps(line))
