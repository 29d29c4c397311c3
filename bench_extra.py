"""
bench_extra.py -- the two remaining BASELINE.json configurations, run through `bench.py --workload S3|S4`.

S3  (configs[2]) layered gll_2_gll on a cubed-sphere spherical shell: order 4, true curved geometry (every GLL node
    on its sphere, equiangular gnomonic projection), coordinates of O(6.4e6) m, three radial layers with a thin crust
    (mantle 2 811 km, lower crust 55 km, upper crust 25 km), one spatial index PER LAYER over the layer's centroids with
    layer-local element ids, as the reference does (components/interpolator.py:363-373); V1 location
    (gll_2_gll_layered, :288-439) and V2 with snap_to_nearest (gll_2_gll_layered_multi_two, :980-1082).  Both meshes are
    generated on the device.  Reported per variant: points/s, per-layer kernel times, candidates Newton-tested per point,
    map evaluations (Newton iterations) per point, failed points; parity of one layer block against the CPU oracle.
S4  (configs[3]) exodus_2_gll (HEX8 nodal source -> order-4 GLL points through the order-1 trilinear path V6,
    components/interpolator.py:142-224, src/trilinearinterpolator.c:40-148) and gll_2_exodus back (V1 over centroids,
    :227-285) with 5 fields + 5 gradient fields; round-trip error against the analytic nodal fields; parity of a sample
    against the oracle.
"""
import ctypes as C
import os
import time

import numpy as np

NAMES = ["QKAPPA", "QMU", "RHO", "VP", "VS"]
UNIT = "points/s"
R_EARTH = 6371000.0
_FACES = ((0, +1.0, 1, 2), (0, -1.0, 2, 1), (1, +1.0, 2, 0), (1, -1.0, 0, 2), (2, +1.0, 0, 1), (2, -1.0, 1, 0))
# (r_bottom, r_top, layer id): mantle, lower crust, thin upper crust -- ids descend with depth like Salvus meshes
SHELL_RADII = ((3480e3, 6291e3, 3), (6291e3, 6346e3, 2), (6346e3, 6371e3, 1))


def shell_mesh_device(n_lat, n_rad, order, dev):
    """meshgen.shell_mesh on the device (torch): coords [E,125,3], layer id per element; elements of one layer are a
    contiguous range (whole shells, bottom shell first)."""
    import torch
    from multimesh_b200.gll import gll_nodes

    z = torch.tensor(gll_nodes(order), dtype=torch.float64, device=dev)
    m = z.numel()
    t = 0.5 * (z + 1.0)
    a = torch.arange(m ** 3, device=dev)
    li, lj, lk = a % m, (a // m) % m, a // (m * m)
    c = torch.arange(n_lat * n_lat, device=dev)
    cu, cv = c % n_lat, c // n_lat
    ncell = 6 * n_lat * n_lat
    dirs = torch.empty((ncell, m ** 3, 3), dtype=torch.float64, device=dev)
    for f, (ax, sgn, a1, a2) in enumerate(_FACES):
        u = (cu[:, None].to(torch.float64) + t[li][None, :]) / n_lat * 2.0 - 1.0
        v = (cv[:, None].to(torch.float64) + t[lj][None, :]) / n_lat * 2.0 - 1.0
        ta, tb = torch.tan(u * (np.pi / 4.0)), torch.tan(v * (np.pi / 4.0))
        vec = dirs[f * n_lat * n_lat:(f + 1) * n_lat * n_lat]
        vec[..., ax] = sgn
        vec[..., a1] = ta
        vec[..., a2] = tb * sgn
        vec /= torch.linalg.norm(vec, dim=-1, keepdim=True)
    nshell = int(sum(n_rad))
    coords = torch.empty((ncell * nshell, m ** 3, 3), dtype=torch.float64, device=dev)
    layer = torch.empty((ncell * nshell,), dtype=torch.int32, device=dev)
    s = 0
    bounds = []
    for (r0, r1, lid), nr in zip(SHELL_RADII, n_rad):
        e0 = s * ncell
        for ir in range(nr):
            rb = r0 + (r1 - r0) * ir / nr
            rt = r0 + (r1 - r0) * (ir + 1) / nr
            r = rb + (rt - rb) * t[lk]
            coords[s * ncell:(s + 1) * ncell] = dirs * r[None, :, None]
            layer[s * ncell:(s + 1) * ncell] = lid
            s += 1
        bounds.append((lid, e0, s * ncell))
    return coords, layer, bounds


def shell_fields_device(coords):
    """Five smooth fields of (radius, direction), [E,5,P]."""
    import torch

    x, y, z = coords[..., 0], coords[..., 1], coords[..., 2]
    r = torch.sqrt(x * x + y * y + z * z)
    d = 1.0 - r / R_EARTH
    out = torch.empty((coords.shape[0], 5, coords.shape[1]), dtype=torch.float64, device=coords.device)
    out[:, 3, :] = 5800.0 + 6000.0 * d + 150.0 * torch.sin(3.0 * x / r) * torch.cos(2.0 * y / r)
    out[:, 4, :] = out[:, 3, :] / np.sqrt(3.0)
    out[:, 2, :] = 2600.0 + 3000.0 * d + 50.0 * (z / r)
    out[:, 0, :] = 57823.0 + 100.0 * torch.cos(np.pi * (x + y + z) / r)
    out[:, 1, :] = 600.0 - 200.0 * d + 20.0 * (x / r) * (y / r)
    return out


def _events():
    import torch

    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def run_shell(args, w, lib, ops, world, rank, dev):
    import torch
    from bench import ClockSampler, full_affinity, hbm_peak, launches_per_step, measure_pipeline, workload_name
    from multimesh_b200 import _lib

    order, k = w["order"], w["k"]
    P, F = 125, 5
    t0 = time.perf_counter()
    src, src_layer, src_bounds = shell_mesh_device(w["n_lat"], w["rad"], order, dev)
    fields = shell_fields_device(src)
    tgt, tgt_layer, tgt_bounds = shell_mesh_device(w["tgt_lat"], w["tgt_rad"], order, dev)
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t0
    E, Et = src.shape[0], tgt.shape[0]
    # per layer: source view + geometry + index over the layer's centroids (layer-local ids), target points
    layers = []
    t0 = time.perf_counter()
    for (lid, e0, e1), (_, f0, f1) in zip(src_bounds, tgt_bounds):
        nodes_l, fields_l = src[e0:e1], fields[e0:e1]
        cent, box = ops.element_geometry(nodes_l)
        pre = ops.element_presolve(nodes_l)
        index = ops.GridIndex(cent)
        pts_l = tgt[f0:f1].reshape(-1, 3)
        # this rank's share of the layer's target points (contiguous range: compact region of the shell)
        n = pts_l.shape[0]
        a, b = (n * rank) // world, (n * (rank + 1)) // world
        layers.append(dict(id=lid, nodes=nodes_l, fields=fields_l, cent=cent, box=box, pre=pre, index=index,
                           pts=pts_l[a:b], n_total=n, info=index.info(), src_range=(e0, e1)))
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    n_local = sum(L["pts"].shape[0] for L in layers)
    n_total = sum(L["n_total"] for L in layers)
    stats_d = torch.zeros(2, dtype=torch.int64, device=dev)
    peak, peak_src = hbm_peak()
    variants = {}
    last = {}
    for vname, spec, kk in (("V1 (gll_2_gll_layered)", ops.V1(), k),
                            ("V2 snap, k=30, tol 1.05 (gll_2_gll_layered_multi_two)", ops.V2(1.05, True), 30)):
        def step():
            res = []
            for L in layers:
                res.append(ops.interpolate(L["index"], 1, L["nodes"], L["cent"], L["box"], L["fields"], L["pts"], kk,
                                           spec, want_location=True, presolve=L["pre"]))
            return res

        sampler = ClockSampler(dev.index)
        sampler.start()
        ms, _, res = measure_pipeline(lib, ops, step, args.steps, args.warmup, world, dev, sampler)
        clocks = sampler.stop()
        # per-layer stage times (one profiled call per layer) and locate statistics (separate instantiation, untimed)
        per_layer = []
        for L, r in zip(layers, res):
            prof = C.c_void_p()
            _lib.check(lib.mm_profile_create(C.byref(prof), 1), "mm_profile_create")
            lib.mm_profile_begin(prof)
            ops.interpolate(L["index"], 1, L["nodes"], L["cent"], L["box"], L["fields"], L["pts"], kk, spec,
                            want_location=True, presolve=L["pre"])
            lib.mm_profile_end()
            nc = C.c_int(0)
            sm = (C.c_float * 6)()
            _lib.check(lib.mm_profile_read(prof, C.byref(nc), sm), "mm_profile_read")
            lib.mm_profile_destroy(prof)
            stats_d.zero_()
            lib.mm_locate_set_stats(C.c_void_p(stats_d.data_ptr()))
            ops.interpolate(L["index"], 1, L["nodes"], L["cent"], L["box"], L["fields"], L["pts"], kk, spec,
                            want_location=True, presolve=L["pre"])
            torch.cuda.synchronize()
            lib.mm_locate_set_stats(None)
            cand, evals = (int(v) for v in stats_d.cpu().tolist())
            npts = L["pts"].shape[0]
            st = torch.bincount(r[3].to(torch.int64), minlength=10).cpu().tolist()
            per_layer.append({
                "layer": L["id"], "source_elements": int(L["nodes"].shape[0]), "points": int(npts),
                "index_cells": L["info"]["cells"], "index_cell_size_m": L["info"]["cell_size"],
                "ms": {"query_sort": sm[0], "K1_knn": sm[1], "K2_locate": sm[2], "rerun": sm[3], "K3_interp": sm[4]},
                "candidates_tested_per_point": cand / max(npts, 1), "map_evaluations_per_point": evals / max(npts, 1),
                "nfailed": int(r[4].item()), "status_histogram": st})
        variants[vname] = {"ms_per_step": ms, "value": n_total / (ms * 1e-3), "unit": UNIT, "per_layer": per_layer,
                           "clocks": clocks}
        if vname.startswith("V1"):
            last[vname] = res  # checked against the oracle below
        del res
    # parity of the V1 result on a block of the thin upper crust (rank 0): the layer's first cube face, a patch of
    # lateral cells, all radial shells of the layer
    parity = None
    if rank == 0 and not args.no_parity:
        from oracle import capi as oracle

        with full_affinity():
            L = layers[-1]
            res = last["V1 (gll_2_gll_layered)"][-1]
            n_lat = w["n_lat"]
            nr = w["rad"][-1]
            ncell = 6 * n_lat * n_lat
            pa = min(n_lat, 20)
            cu = torch.arange(pa, device=dev)
            cells = (cu[None, :] + n_lat * cu[:, None]).reshape(-1)  # face 0, patch [0,pa)^2
            ids = (cells[None, :] + ncell * torch.arange(nr, device=dev)[:, None]).reshape(-1)  # layer-local ids
            sub_nodes = L["nodes"][ids].cpu().numpy()
            sub_fields = L["fields"][ids].cpu().numpy()
            # target points well inside the patch: direction within the patch's angular range shrunk by 3 cells
            p = L["pts"]
            u = torch.atan2(p[:, 1], p[:, 0]) / (np.pi / 4.0)  # face 0: x = +r, (y, z) tangent
            v = torch.atan2(p[:, 2], p[:, 0]) / (np.pi / 4.0)
            lo, hi = -1.0 + 2.0 * 3 / n_lat, -1.0 + 2.0 * (pa - 3) / n_lat
            inside = ((p[:, 0] > 0) & (u > lo) & (u < hi) & (v > lo) & (v < hi)).nonzero().reshape(-1)[:100_000]
            if inside.numel():
                pts = p[inside].cpu().numpy()
                cands = oracle.knn_ckdtree_canonical(oracle.centroids(sub_nodes), pts, k, pad=12)
                o_elem, o_xi, _, _ = oracle.locate(order, 3, sub_nodes, pts, cands, oracle.V1())
                o_out = oracle.interp(order, 3, sub_fields, o_elem, o_xi)
                ids_h = ids.cpu().numpy()
                o_glob = np.where(o_elem >= 0, ids_h[np.maximum(o_elem, 0)], -1)
                g_elem = res[1][inside].cpu().numpy()
                g_out = res[0][inside].cpu().numpy()
                eq = bool(np.array_equal(g_elem, o_glob))
                mr = float(np.max(np.abs(g_out - o_out) / np.maximum(np.abs(o_out), 1e-300)))
                parity = {"points": int(len(pts)), "elem_equal": eq, "max_rel": mr, "tolerance_rel": 1e-10,
                          "what": f"V1 result of the upper-crust layer vs the CPU oracle on a {pa}x{pa} patch of face 0"}
                assert eq and mr <= 1e-10, parity
    v1 = variants["V1 (gll_2_gll_layered)"]
    return {
        "metric": "target GLL points interpolated/sec", "value": v1["value"], "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": v1["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(w), "source_elements": int(E), "target_elements": int(Et),
                   "points_total": int(n_total), "points_this_rank": int(n_local), "fields": F,
                   "source_gb": round((src.numel() + fields.numel()) * 8 / 1e9, 1),
                   "l2": "inputs larger than L2, no flush"},
        "variants": variants, "parity_check": parity, "mesh_generation_s": gen_s, "geometry_index_build_s": build_s,
        "roofline": {"bound": "hbm", "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                     "note": "per-layer kernel times in variants[*].per_layer; K3 algorithmic bytes 5 068 B/point"},
        "cpu_baseline": None, "e2e": None, "clocks": v1["clocks"],
        # V1 variant: one mm_interpolate per layer and step, centroid form
        "gpu_launches": int(sum(launches_per_step(int(L["pts"].shape[0]), True, False) for L in layers) * args.steps),
    }


def run_exodus(args, w, lib, ops, world, rank, dev):
    import torch
    from bench import ClockSampler, full_affinity, hbm_peak, launches_per_step, rerun_rounds, workload_name
    from multimesh_b200 import meshgen

    order, k = w["order"], w["k"]
    names = NAMES + ["grad" + n for n in NAMES]
    Fn = len(names)
    points, conn = meshgen.hex8_mesh((w["hex"],) * 3, warp=0.01)
    gll = meshgen.box_mesh((w["gll"],) * 3, order, lo=[0.02] * 3, hi=[0.98] * 3, warp=0.01)
    x, y, z = points[:, 0], points[:, 1], points[:, 2]
    base = [57823.0 + 100.0 * np.cos(np.pi * (x + y + z)), 600.0 - 80.0 * x + 40.0 * y * z,
            2600.0 + 300.0 * x * y + 150.0 * z * z, 5000.0 + 800.0 * np.sin(2 * np.pi * x) * np.cos(2 * np.pi * y) + 300 * z,
            (5000.0 + 800.0 * np.sin(2 * np.pi * x) * np.cos(2 * np.pi * y) + 300 * z) / np.sqrt(3.0)]
    nodal = np.stack(base + [1e-3 * b * (1.0 + x) for b in base])  # [10, Np]
    perm = np.argsort([0, 3, 2, 1, 4, 5, 6, 7])  # Exodus -> vertex order of the C routine (interpolator.py:186-190)
    connC = np.ascontiguousarray(conn[:, perm])
    t_points, t_conn, t_connC = (torch.from_numpy(a).to(dev) for a in (points, conn, connC))
    t_nodal = torch.from_numpy(nodal).to(dev)
    t_gll = torch.from_numpy(gll).to(dev)
    gpts = t_gll.view(-1, 3)
    N1 = gpts.shape[0]
    cent8 = ops.centroid_conn(t_conn, t_points)
    index8 = ops.GridIndex(cent8)
    ev = {n: _events() for n in ("search", "gather", "knn", "tri")}
    res = {}

    def exodus_2_gll():
        # centroid_tree.query(k) + triLinearInterpolator as the one progressive call the driver makes
        # (components/interpolator.py: exodus_2_gll), then the gather of the 10 nodal fields
        ev["search"][0].record()
        nf, enc, wts = ops.trilinear_indexed(index8, t_connC, t_points, gpts, k)
        ev["search"][1].record()
        ev["gather"][0].record()
        vals = ops.gather_nodal(t_nodal, enc, wts)  # [F, N]
        ev["gather"][1].record()
        res.update(nf=nf, enc=enc, wts=wts, vals=vals)
        return vals

    def two_step():
        # the reference's two calls one after the other (complete k-NN lists, then the C routine's twin): the
        # progressive call has to reproduce this bit for bit
        ev["knn"][0].record()
        nn = index8.query_idx(gpts, k)
        ev["knn"][1].record()
        ev["tri"][0].record()
        nf, enc, wts = ops.trilinear(nn.to(torch.int64), t_connC, t_points, gpts)
        ev["tri"][1].record()
        return nn, nf, enc, wts

    sampler = ClockSampler(dev.index)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        exodus_2_gll()
    torch.cuda.synchronize()
    e0, e1 = _events()
    w0 = time.time()
    e0.record()
    for _ in range(args.steps):
        exodus_2_gll()
    e1.record()
    torch.cuda.synchronize()
    sampler.mark(w0, time.time(), "timed region")
    ms1 = e0.elapsed_time(e1) / args.steps
    assert int(res["nf"].item()) == 0
    two_step()
    nn2, nf2, enc2, wts2 = two_step()
    torch.cuda.synchronize()
    stage1 = {n: a.elapsed_time(b) for n, (a, b) in ev.items()}
    same_as_two_step = bool(int(nf2.item()) == 0 and torch.equal(enc2, res["enc"]) and torch.equal(wts2, res["wts"]))
    assert same_as_two_step, "mm_trilinear_indexed differs from mm_knn + mm_trilinear"
    res["nn"] = nn2
    del nf2, enc2, wts2
    # [F, N] -> MODEL/data layout [E, F, P]
    gll_data = res["vals"].view(Fn, gll.shape[0], 125).permute(1, 0, 2).contiguous()
    # ---- gll_2_exodus: V1 over centroids, all 10 fields, targets = the exodus nodes inside the GLL mesh
    inside = np.all((points > 0.03) & (points < 0.97), axis=1)
    t_back = torch.from_numpy(np.ascontiguousarray(points[inside])).to(dev)
    N2 = t_back.shape[0]
    cent, box = ops.element_geometry(t_gll)
    pre = ops.element_presolve(t_gll)
    index = ops.GridIndex(cent)

    def gll_2_exodus():
        return ops.interpolate(index, 1, t_gll, cent, box, gll_data, t_back, k, ops.V1(), presolve=pre)

    for _ in range(max(args.warmup, 3)):
        back = gll_2_exodus()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        gll_2_exodus()
    e1.record()
    torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / args.steps
    clocks = sampler.stop()
    back_vals = back[0].cpu().numpy()  # [N2, F]
    exact = nodal[:, inside].T
    rt_err = float(np.max(np.abs(back_vals - exact) / np.abs(exact)))
    # ---- parity of a sample of the V6 path against the oracle (and, where built, the compiled reference C)
    parity = None
    if rank == 0 and not args.no_parity:
        from oracle import capi as oracle

        with full_affinity():
            sel = np.arange(0, N1, max(1, N1 // 200_000))
            q = gll.reshape(-1, 3)[sel]
            nn_h = res["nn"][torch.from_numpy(sel).to(dev)].cpu().numpy().astype(np.int64)
            o_nf, o_enc, o_w = oracle.trilinear_interpolator(k, nn_h, connC, points, q)
            g_enc = res["enc"][torch.from_numpy(sel).to(dev)].cpu().numpy()
            g_w = res["wts"][torch.from_numpy(sel).to(dev)].cpu().numpy()
            parity = {"points": int(len(sel)), "enclosing_equal": bool(np.array_equal(g_enc, o_enc)),
                      "weights_bit_equal": bool(np.array_equal(g_w, o_w)), "oracle_failed": int(o_nf),
                      "what": "exodus_2_gll: enclosing node ids and trilinear weights vs the CPU oracle on a sample"}
            if oracle.ref_lib() is not None:
                r_nf, r_enc, r_w = oracle.ref_trilinear_interpolator(k, nn_h[:20000], connC, points, q[:20000])
                parity["reference_c_bit_equal"] = bool(np.array_equal(g_enc[:20000], r_enc) and
                                                       np.array_equal(g_w[:20000], r_w))
            assert parity["enclosing_equal"] and parity["weights_bit_equal"], parity
    peak, peak_src = hbm_peak()
    total_pts = N1 + N2
    total_ms = ms1 + ms2
    return {
        "metric": "target GLL points interpolated/sec", "value": total_pts / (total_ms * 1e-3), "unit": UNIT,
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(w), "hex8_elements": int(conn.shape[0]), "gll_elements": int(gll.shape[0]),
                   "fields": Fn, "l2": "exodus_2_gll reads 8 M target points + 2.1 M-element source per step, > L2"},
        "exodus_2_gll": {"points": int(N1), "ms_per_step": ms1, "value": N1 / (ms1 * 1e-3), "unit": UNIT,
                         "ms": {"search_progressive_k20_V6_trilinear": stage1["search"],
                                "gather_nodal_10_fields": stage1["gather"]},
                         "two_step_ms": {"K1_knn_k20": stage1["knn"], "V6_trilinear": stage1["tri"]},
                         "equals_two_step": same_as_two_step, "nfailed": 0},
        "gll_2_exodus": {"points": int(N2), "ms_per_step": ms2, "value": N2 / (ms2 * 1e-3), "unit": UNIT,
                         "nfailed": int(back[4].item())},
        "round_trip_max_rel_error": rt_err,
        "round_trip_note": "nodal field -> order-4 GLL (trilinear) -> back to the nodes (order-4 Lagrange): the error "
                           "is the trilinear interpolation error of the smooth fields on the 128^3 HEX8 mesh",
        "parity_check": parity,
        "roofline": {"bound": "hbm", "peak": peak, "peak_source": peak_src, "unit": "GB/s"},
        "cpu_baseline": None, "e2e": None, "clocks": clocks,
        # exodus_2_gll: query sort 5, tile first pass 1, prefix search 1, 3 per re-run round, nodal gather 1;
        # gll_2_exodus: one mm_interpolate, centroid form
        "gpu_launches": int((8 + 3 * rerun_rounds(N1) + launches_per_step(N2, True, False)) * args.steps),
    }


def run_quads(args, w, lib, ops, world, rank, dev):
    """S1 (BASELINE configs[0], the reference's own CPU-runnable case): 2-D gll_2_gll on quads -- source 256 x 256
    elements of order 2 (P = 9; `order=4`, P = 25, is the other 2-D order the reference supports) on [0,1]^2 with
    VP, VS, RHO, targets = the GLL points of a non-nested 200 x 200 mesh, GLL-point k-NN form, V1 location.  The whole
    target set is also run through the CPU oracle (parity of every point) and timed there (`cpu_baseline`)."""
    import torch
    from bench import ClockSampler, full_affinity, hbm_peak, launches_per_step, measure_pipeline, workload_name
    from multimesh_b200 import meshgen

    order, k = w["order"], w["k"]
    names = ["VP", "VS", "RHO"]
    P, F = (order + 1) ** 2, len(names)
    nodes_h = meshgen.box_mesh((w["src"],) * 2, order, warp=0.01)
    x, y = nodes_h[..., 0], nodes_h[..., 1]
    vp = 5000.0 + 800.0 * np.sin(2 * np.pi * x) * np.cos(2 * np.pi * y)
    fields_h = np.ascontiguousarray(np.stack([vp, vp / np.sqrt(3.0), 2600.0 + 300.0 * x * y], axis=1))
    shift = (rank % 8) * 0.11 / w["tgt"]
    pts_h = np.ascontiguousarray(meshgen.box_mesh((w["tgt"],) * 2, order, lo=[0.001 + 0.1 * shift] * 2,
                                                  hi=[0.999 - shift] * 2).reshape(-1, 2))
    E, N = nodes_h.shape[0], pts_h.shape[0]
    nodes, fields, pts = (torch.from_numpy(a).to(dev) for a in (nodes_h, fields_h, pts_h))
    t0 = time.perf_counter()
    cent, box = ops.element_geometry(nodes)
    pre = ops.element_presolve(nodes)
    index = ops.GridIndex(nodes.view(E * P, 2)).prepare_sites()
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0

    def step():
        return ops.interpolate(index, P, nodes, cent, box, fields, pts, k, ops.V1(), want_location=True, presolve=pre)

    sampler = ClockSampler(dev.index)
    sampler.start()
    ms, stages, res = measure_pipeline(lib, ops, step, args.steps, args.warmup, world, dev, sampler)
    clocks = sampler.stop()
    out, elem, xi, status, nfail = res
    peak, peak_src = hbm_peak()
    bytes_pt = {"K1_knn": 8 * 2 + 4 * 8, "K2_locate": 8 * 2 + (8 * 2 * P + 16 * 2) + (4 + 8 * 2),
                "K3_interp": (8 * 2 + 4) + 8 * F * P + 8 * F}
    kernels = {}
    for name, t in (("K1_knn", stages[1]), ("K2_locate", stages[2]), ("K3_interp", stages[4])):
        gbs = bytes_pt[name] * N / (t * 1e-3) / 1e9
        kernels[name] = {"ms": round(float(t), 4), "alg_bytes_per_point": bytes_pt[name], "achieved_gbs": round(gbs, 1),
                         "frac": round(gbs / peak, 4)}
    parity = cpu = None
    if rank == 0 and not args.no_parity:
        from oracle import capi as oracle

        with full_affinity():
            oracle.set_num_threads(len(os.sched_getaffinity(0)))
            t0 = time.perf_counter()
            cands = (oracle.knn_ckdtree_canonical(nodes_h.reshape(-1, 2), pts_h, k, pad=24) // P).astype(np.int32)
            t_knn = time.perf_counter() - t0
            t0 = time.perf_counter()
            o_elem, o_xi, o_st, o_nf = oracle.locate(order, 2, nodes_h, pts_h, cands, oracle.V1())
            o_out = oracle.interp(order, 2, fields_h, o_elem, o_xi)
            t_rest = time.perf_counter() - t0
        g_out, g_elem = out.cpu().numpy(), elem.cpu().numpy()
        eq = bool(np.array_equal(g_elem, o_elem))
        mr = float(np.max(np.abs(g_out - o_out) / np.maximum(np.abs(o_out), 1e-300)))
        parity = {"points": int(N), "elem_equal": eq, "max_rel": mr, "bit_equal_values": bool(np.array_equal(g_out, o_out)),
                  "xi_bit_equal": bool(np.array_equal(xi.cpu().numpy(), o_xi)), "tolerance_rel": 1e-10,
                  "what": "EVERY target point of the timed run vs the CPU oracle (canonical k-NN, C locate / gather)"}
        assert eq and mr <= 1e-10 and int(nfail.item()) == o_nf, parity
        if world == 1 and not args.no_cpu:
            cpu = {"value": N / (t_knn + t_rest), "unit": UNIT, "cores": int(oracle.num_threads()), "kind": "port",
                   "sample": f"all {N} target points once: cKDTree over-query + canonical re-sort {t_knn:.2f} s, C "
                             f"locate + gather {t_rest:.2f} s (tree build included)"}
    dom = max(kernels, key=lambda n: kernels[n]["ms"])
    return {
        "metric": "target GLL points interpolated/sec", "value": world * N / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(w), "points_per_gpu": int(N), "source_elements": int(E), "fields": F,
                   "l2": "inputs (9.5 MB source, 5.8 MB targets) fit the 126 MB L2: steps after the first run from L2 -- "
                         "this configuration is the reference's laptop-sized case, latency- not bandwidth-bound"},
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                     "frac": kernels[dom]["frac"], "traffic": None, "peak_source": peak_src},
        "kernels": kernels,
        "other_stages": {"query_sort_ms": round(float(stages[0]), 4), "rerun_unresolved_ms": round(float(stages[3]), 4)},
        "parity_check": parity, "cpu_baseline": cpu, "e2e": None, "clocks": clocks, "geometry_index_build_s": build_s,
        "nfailed": int(nfail.item()), "status_histogram": torch.bincount(status.to(torch.int64), minlength=10).cpu().tolist(),
        "gpu_launches": int(launches_per_step(N, True, True) * args.steps),
    }


def run(args, w, lib, ops, world, rank, dev, all_cpus):
    if w["kind"] == "shell":
        return run_shell(args, w, lib, ops, world, rank, dev)
    if w["kind"] == "quads":
        return run_quads(args, w, lib, ops, world, rank, dev)
    return run_exodus(args, w, lib, ops, world, rank, dev)
