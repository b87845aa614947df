"""Seeded synthetic designs in the reference's batch format (host side, numpy).

The reference ships no data; its batches come from ``generate_data.py:50-54``
(a 7-tuple: heterograph, topo_levels, path_masks, path2level, path2endpoint,
critical_paths, cnn_inputs).  ``make_design`` builds the same information for
a random levelised netlist:

* ``n_pi = max(64, n_cells//50)`` primary-input pins on level 0;
* ``n_lv`` cell-levels of ``n_cells//n_lv`` cells; fan-in k ~ {1:.2,2:.4,3:.3,4:.1};
  input pin 0 of a cell is driven from the previous cell-level, the others by
  any earlier driver, so edges skip levels;
* one ``net`` edge driver->sink per input pin, one ``cell`` edge sink->output;
* ``cell_feat (N,36)``: one-hot(34) + 2 uniforms on output/PI pins, zero on sinks;
  ``net_feat (N,2)``: uniforms on sinks;
* pin bins: a cell sits within ``spread`` bins of the driver of its pin 0
  (nets are local, as in placed designs, so path masks stay sparse);
* ``n_endpoints`` sink pins as timing endpoints, ``arrival_time ~ U[0,1)``;
* path masks: critical-path trace-back + union of bin bounding boxes
  (``verilog_parser_asap7.py:1433-1450,1315-1351``), CSR over ``map*map`` columns;
* image ``(C,H,H) ~ U[0,1)`` with ``H = 2*map`` for the UNet (``Unet.py:74-78``
  halves the resolution) or ``H = 4*map`` for LayoutNet.

Pin ids are shuffled (netlist order is unrelated to level order in real
designs), so level rows are scattered in HBM exactly as they would be.
"""
from dataclasses import dataclass, field

import numpy as np


@dataclass
class SynthDesign:
    n: int
    net_src: np.ndarray
    net_dst: np.ndarray
    cell_src: np.ndarray
    cell_dst: np.ndarray
    cell_feat: np.ndarray
    net_feat: np.ndarray
    pis: np.ndarray
    level: np.ndarray            # pin -> level (int32)
    pin_xy: np.ndarray           # (n,2) int32 bins
    endpoints: np.ndarray        # (P,) int64 pin ids, path id = position
    arrival_time: np.ndarray     # (P,) fp32
    mask_indptr: np.ndarray      # (P+1,) int32
    mask_cols: np.ndarray        # (nnz,) int32, ascending per row
    image: np.ndarray            # (C,H,H) fp32
    map_size: int
    meta: dict = field(default_factory=dict)

    @property
    def num_levels(self):
        return int(self.level.max()) + 1

    def level_lists(self):
        """Per level: ascending pin ids (the reference's in-level order is undefined)."""
        order = np.argsort(self.level, kind="stable")
        order = order[self.level[order] >= 0]
        counts = np.bincount(self.level[order], minlength=self.num_levels)
        return np.split(order, np.cumsum(counts)[:-1])

    def topo_levels(self):
        """``topo_levels`` as ``cal_topo_level`` returns them
        (verilog_parser_asap7.py:1505-1508): [(nodes, targets, path_ids)] Python lists."""
        ep_level = self.level[self.endpoints]
        out = []
        for lid, nodes in enumerate(self.level_lists()):
            pid = np.nonzero(ep_level == lid)[0]
            out.append((nodes.tolist(), self.endpoints[pid].tolist(), pid.tolist()))
        return out

    def path_dicts(self):
        """``path2level`` / ``path2endpoint`` (dataset.py:106-131)."""
        ep_level = self.level[self.endpoints]
        p2l = {i: int(l) for i, l in enumerate(ep_level)}
        p2e = {i: int(e) for i, e in enumerate(self.endpoints)}
        return p2l, p2e


def longest_path_levels(n, src, dst, pis):
    """pin -> level by Kahn peeling over the part reachable from ``pis``; -1 elsewhere."""
    src = np.asarray(src, np.int64)
    dst = np.asarray(dst, np.int64)
    order = np.argsort(src, kind="stable")
    optr = np.zeros(n + 1, np.int64)
    np.cumsum(np.bincount(src, minlength=n), out=optr[1:])
    odst = dst[order]

    def fan_out(nodes):
        cnt = optr[nodes + 1] - optr[nodes]
        tot = int(cnt.sum())
        if tot == 0:
            return np.zeros(0, np.int64)
        off = np.arange(tot) - np.repeat(np.cumsum(cnt) - cnt, cnt)
        return odst[np.repeat(optr[nodes], cnt) + off]

    reach = np.zeros(n, bool)
    fr = np.unique(np.asarray(pis, np.int64))
    reach[fr] = True
    while fr.size:
        nb = np.unique(fan_out(fr))
        nb = nb[~reach[nb]]
        reach[nb] = True
        fr = nb
    indeg = np.bincount(dst[reach[src] & reach[dst]], minlength=n)
    level = np.full(n, -1, np.int32)
    fr = np.nonzero(reach & (indeg == 0))[0]
    k = 0
    while fr.size:
        level[fr] = k
        nb = fan_out(fr)
        if nb.size:
            indeg -= np.bincount(nb, minlength=n)
            cand = np.unique(nb)
            fr = cand[indeg[cand] == 0]
        else:
            fr = nb
        k += 1
    return level


def _trace_masks(n, src, dst, level, endpoints, pin_xy, map_size):
    """Vectorised critical-path trace + bbox rasterisation -> CSR (indptr, cols)."""
    order = np.argsort(dst, kind="stable")                   # in-edges in edge-id order
    iptr = np.zeros(n + 1, np.int64)
    np.cumsum(np.bincount(dst, minlength=n), out=iptr[1:])
    isrc = src[order]
    maxdeg = int((iptr[1:] - iptr[:-1]).max()) if n else 0
    P = endpoints.size
    dense = np.zeros((P, map_size, map_size), bool)
    cur = endpoints.astype(np.int64).copy()
    cur_level = level[cur].astype(np.int64)
    active = cur_level >= 2
    while active.any():
        idx = np.nonzero(active)[0]
        c = cur[idx]
        nxt = np.full(idx.size, -1, np.int64)
        deg = iptr[c + 1] - iptr[c]
        for j in range(maxdeg):
            ok = (deg > j) & (nxt < 0)
            cand = isrc[np.minimum(iptr[c] + j, isrc.size - 1)]
            ok &= level[cand] == cur_level[idx] - 1
            nxt[ok] = cand[ok]
        if (nxt < 0).any():
            raise RuntimeError("pin without a predecessor one level below")
        a, b = pin_xy[c], pin_xy[nxt]
        x1, x2 = np.minimum(a[:, 0], b[:, 0]), np.maximum(a[:, 0], b[:, 0])
        y1, y2 = np.minimum(a[:, 1], b[:, 1]), np.maximum(a[:, 1], b[:, 1])
        for p, xa, xb, ya, yb in zip(idx.tolist(), x1.tolist(), x2.tolist(), y1.tolist(), y2.tolist()):
            dense[p, xa:xb + 1, ya:yb + 1] = True
        cur[idx] = nxt
        cur_level[idx] -= 1
        active = cur_level >= 2
    flat = dense.reshape(P, -1)
    indptr = np.zeros(P + 1, np.int32)
    np.cumsum(flat.sum(1), out=indptr[1:])
    cols = np.nonzero(flat)[1].astype(np.int32)
    return indptr, cols


def make_design(n_cells=5000, n_lv=20, map_size=32, n_endpoints=1350, seed=0,
                img_channels=3, img_scale=2, shuffle_ids=True, spread=2, num_ctypes=34):
    rng = np.random.default_rng(seed)
    n_pi = max(64, n_cells // 50)
    cpl = max(1, n_cells // n_lv)
    net_s, net_d, cell_s, cell_d = [], [], [], []
    drivers_all = [np.arange(n_pi, dtype=np.int64)]
    prev = drivers_all[0]
    xy = [rng.integers(0, map_size, size=(n_pi, 2))]
    is_out = [np.ones(n_pi, bool)]
    base = n_pi
    for _ in range(n_lv):
        k = rng.choice(np.array([1, 2, 3, 4]), p=[.2, .4, .3, .1], size=cpl)
        tot = int(k.sum())
        first = np.cumsum(k) - k
        cell_of = np.repeat(np.arange(cpl), k)
        slot = np.arange(tot) - first[cell_of]
        sinks = base + np.arange(tot, dtype=np.int64)
        outs = base + tot + np.arange(cpl, dtype=np.int64)
        earlier = np.concatenate(drivers_all)
        drv = earlier[rng.integers(0, earlier.size, size=tot)]
        d0 = prev[rng.integers(0, prev.size, size=cpl)]
        drv[slot == 0] = d0
        net_s.append(drv); net_d.append(sinks)
        cell_s.append(sinks); cell_d.append(outs[cell_of])
        allxy = np.concatenate(xy)
        cxy = np.clip(allxy[d0] + rng.integers(-spread, spread + 1, size=(cpl, 2)), 0, map_size - 1)
        xy.append(cxy[cell_of]); xy.append(cxy)
        is_out.append(np.zeros(tot, bool)); is_out.append(np.ones(cpl, bool))
        drivers_all.append(outs)
        prev = outs
        base += tot + cpl
    n = base
    net_src, net_dst = np.concatenate(net_s), np.concatenate(net_d)
    cell_src, cell_dst = np.concatenate(cell_s), np.concatenate(cell_d)
    pin_xy = np.concatenate(xy).astype(np.int32)
    is_out = np.concatenate(is_out)
    pis = np.arange(n_pi, dtype=np.int64)
    if shuffle_ids:
        perm = rng.permutation(n)                            # old id -> new id
        net_src, net_dst, cell_src, cell_dst = perm[net_src], perm[net_dst], perm[cell_src], perm[cell_dst]
        pis = perm[pis]
        inv = np.empty(n, np.int64); inv[perm] = np.arange(n)
        pin_xy, is_out = pin_xy[inv], is_out[inv]
    cell_feat = np.zeros((n, num_ctypes + 2), np.float32)
    oi = np.nonzero(is_out)[0]
    cell_feat[oi, rng.integers(0, num_ctypes, size=oi.size)] = 1.0
    cell_feat[oi, num_ctypes:] = rng.random((oi.size, 2), dtype=np.float32)
    net_feat = np.zeros((n, 2), np.float32)
    si = np.nonzero(~is_out)[0]
    net_feat[si] = rng.random((si.size, 2), dtype=np.float32)
    src = np.concatenate([net_src, cell_src]); dst = np.concatenate([net_dst, cell_dst])
    level = longest_path_levels(n, src, dst, pis)
    P = min(n_endpoints, si.size)
    endpoints = np.sort(rng.choice(si, size=P, replace=False)).astype(np.int64)
    endpoints = endpoints[np.argsort(level[endpoints], kind="stable")]   # grouped by level
    arrival = rng.random(P, dtype=np.float32)
    mi, mc = _trace_masks(n, src, dst, level, endpoints, pin_xy, map_size)
    H = img_scale * map_size
    image = rng.random((img_channels, H, H), dtype=np.float32)
    return SynthDesign(n=n, net_src=net_src, net_dst=net_dst, cell_src=cell_src, cell_dst=cell_dst,
                       cell_feat=cell_feat, net_feat=net_feat, pis=pis, level=level, pin_xy=pin_xy,
                       endpoints=endpoints, arrival_time=arrival, mask_indptr=mi, mask_cols=mc,
                       image=image, map_size=map_size,
                       meta=dict(n_cells=n_cells, n_lv=n_lv, seed=seed, n_pi=n_pi))


CONFIGS = {
    # BASELINE.json configs
    "c1": dict(n_cells=5000, n_lv=20, map_size=32),          # 64x64 image, CPU reference point
    "c2": dict(n_cells=100000, n_lv=50, map_size=128),       # 256x256 image, headline
    "c3": dict(n_cells=300000, n_lv=50, map_size=128),       # GNN-only, ~1M pins
    "tiny": dict(n_cells=300, n_lv=6, map_size=8, n_endpoints=40),
}
