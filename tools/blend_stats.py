#!/usr/bin/env python
"""Blend-kernel work statistics of a synthetic scene, computed on the CPU in numpy (float64; a design
aid, not a parity tool): for a sample of tiles, which (pixel, Gaussian) pairs are blended, and how many
(pixel block, entry) hits / candidate visits different block shapes and list-walking schemes of the
backward blend kernel would execute.  Usage: python tools/blend_stats.py [--config C3] [--every 8]"""
import argparse, math, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge


def geometry(cam, scene):
    m = scene.means3D.double().numpy(); s = scene.scales.double().numpy(); q = scene.rotations.double().numpy()
    o = scene.opacities.double().numpy()[:, 0]
    W2C = cam.w2c.double().numpy(); R, T = W2C[:3, :3], W2C[:3, 3]
    t = m @ R.T + T
    fx, fy = cam.W / (2 * cam.tanfovx), cam.H / (2 * cam.tanfovy)
    r, x, y, z = q.T
    Rm = np.stack([np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)], -1),
                   np.stack([2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)], -1),
                   np.stack([2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], -1)], 1)
    Sig = Rm @ (s[:, :, None] ** 2 * np.transpose(Rm, (0, 2, 1)))
    tz = t[:, 2]
    lx, ly = 1.3 * cam.tanfovx, 1.3 * cam.tanfovy
    tx = np.clip(t[:, 0] / tz, -lx, lx) * tz; ty = np.clip(t[:, 1] / tz, -ly, ly) * tz
    J = np.zeros((len(m), 2, 3)); J[:, 0, 0] = fx / tz; J[:, 0, 2] = -fx * tx / tz ** 2; J[:, 1, 1] = fy / tz; J[:, 1, 2] = -fy * ty / tz ** 2
    A = J @ R
    cov = A @ Sig @ np.transpose(A, (0, 2, 1))
    a, b, c = cov[:, 0, 0] + 0.3, cov[:, 0, 1], cov[:, 1, 1] + 0.3
    det = a * c - b * b
    cA, cB, cC = c / det, -b / det, a / det
    px = ((t[:, 0] / tz * fx / (cam.W / 2) + 1) * cam.W - 1) * 0.5   # ndc2pix of x_ndc = fx*tx/tz / (W/2)
    py = ((t[:, 1] / tz * fy / (cam.H / 2) + 1) * cam.H - 1) * 0.5
    mid = 0.5 * (a + c); lam = mid + np.sqrt(np.maximum(0.1, mid * mid - det)); rad = np.ceil(3 * np.sqrt(lam))
    vis = (tz > 0.2) & (det != 0)
    with np.errstate(divide="ignore", invalid="ignore"):
        pc = np.where(o > 0, np.log((15 / 255) / o) - 1e-3, 1.0)
    cdet = cA * cC - cB * cB
    k = -2 * pc / cdet
    hx = np.sqrt(np.maximum(k * cC, 0)) * 1.0001 + 0.02; hy = np.sqrt(np.maximum(k * cA, 0)) * 1.0001 + 0.02
    vis &= pc <= 0
    return dict(px=px, py=py, A=cA, B=cB, C=cC, o=o, pc=pc, depth=tz, rad=rad, hx=hx, hy=hy, vis=vis)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C3"); ap.add_argument("--every", type=int, default=8)
    a = ap.parse_args()
    sc = ge.load_scene_module()
    cam, scene = sc.config(a.config)
    g = geometry(cam, scene)
    W, H = cam.W, cam.H
    gx, gy = (W + 15) // 16, (H + 15) // 16
    vis = np.nonzero(g["vis"])[0]
    px, py, hx, hy, rad = (g[k][vis] for k in ("px", "py", "hx", "hy", "rad"))
    # tight tile rectangles (preprocess_fwd: 3-sigma rect intersected with the cut ellipse's box)
    tx0 = np.maximum(np.clip(((px - rad) // 16), 0, gx), np.ceil((px - hx - 15) / 16)).astype(int)
    tx1 = np.minimum(np.clip(((px + rad + 15) // 16), 0, gx), np.floor((px + hx) / 16) + 1).astype(int)
    ty0 = np.maximum(np.clip(((py - rad) // 16), 0, gy), np.ceil((py - hy - 15) / 16)).astype(int)
    ty1 = np.minimum(np.clip(((py + rad + 15) // 16), 0, gy), np.floor((py + hy) / 16) + 1).astype(int)
    tx0, ty0 = np.clip(tx0, 0, gx), np.clip(ty0, 0, gy); tx1, ty1 = np.clip(tx1, 0, gx), np.clip(ty1, 0, gy)
    ntiles = np.maximum(tx1 - tx0, 0) * np.maximum(ty1 - ty0, 0)
    print("visible %d of %d, duplicates N = %d (%.1f per tile)" % (len(vis), len(g["vis"]), ntiles.sum(), ntiles.sum() / (gx * gy)))
    tiles = [(x, y) for y in range(gy) for x in range(gx)][::a.every]
    lx, ly = np.meshgrid(np.arange(16), np.arange(16))
    lx, ly = lx.reshape(-1), ly.reshape(-1)
    blk = {"8x8": (ly // 8) * 2 + lx // 8, "8x4": (ly // 4) * 2 + lx // 8, "4x4": (ly // 4) * 4 + lx // 4,
           "4x2": (ly // 2) * 4 + lx // 4,
           "16x4": ly // 4, "16x8": ly // 8, "16x16": lx * 0}
    tot = dict(pairs=0, entries=0, walk=0)
    hits = {k: 0 for k in blk}; cands = {k: 0 for k in blk}
    paired = {"8x8|2x(8x4) halves: full iterations": 0, "8x8|2x(8x4) halves: all iterations": 0,
              "8x4|2x(4x4) halves: full iterations": 0, "8x4|2x(4x4) halves: all iterations": 0,
              "8x8|4x(4x4) quarters: full iterations": 0, "8x8|4x(4x4) quarters: all iterations": 0,
              "8x4|4x(4x2) quarters (1 px / lane): full iterations": 0, "8x4|4x(4x2) quarters (1 px / lane): all iterations": 0,
              "8x8 hits containing a first contributor": 0,
              "16x8|8x(4x4) eighths (4 px / lane): full iterations": 0, "16x8|8x(4x4) eighths (4 px / lane): all iterations": 0,
              "16x8|4x(8x4) quarters (4 px / lane): full iterations": 0, "16x8|4x(8x4) quarters (4 px / lane): all iterations": 0,
              "8x16|8x(4x4) eighths (4 px / lane): full iterations": 0, "8x16|8x(4x4) eighths (4 px / lane): all iterations": 0}
    hist_valid = np.zeros(65, dtype=np.int64)
    t0 = time.time()
    for (tx, ty) in tiles:
        sel = np.nonzero((tx0 <= tx) & (tx < tx1) & (ty0 <= ty) & (ty < ty1))[0]
        if len(sel) == 0:
            continue
        ids = vis[sel]
        order = np.argsort(g["depth"][ids], kind="stable")
        ids = ids[order]
        X, Y = tx * 16 + lx, ty * 16 + ly
        inside = (X < W) & (Y < H)
        dx = g["px"][ids][:, None] - X[None]; dy = g["py"][ids][:, None] - Y[None]
        A, B, C_, o, pc = (g[k][ids][:, None] for k in ("A", "B", "C", "o", "pc"))
        pw = -0.5 * (A * dx * dx + C_ * dy * dy) - B * dx * dy
        al = np.minimum(0.99, o * np.exp(np.minimum(pw, 0)))
        valid = (pw <= 0) & (pw >= pc) & (al >= 15 / 255) & inside[None]
        Tb = np.cumprod(1 - al * valid, axis=0)
        Tbefore = np.vstack([np.ones((1, 256)), Tb[:-1]])
        contrib = valid & ~(Tbefore < 1e-4)
        used = np.nonzero(contrib.any(1))[0]
        walk = (used[-1] + 1) if len(used) else 0
        contrib = contrib[:walk]
        tot["pairs"] += int(contrib.sum()); tot["entries"] += len(ids); tot["walk"] += int(walk)
        # candidate masks from the cut ellipse's bounding box (what the staging thread computes)
        bx0 = g["px"][ids][:walk] - g["hx"][ids][:walk]; bx1 = g["px"][ids][:walk] + g["hx"][ids][:walk]
        by0 = g["py"][ids][:walk] - g["hy"][ids][:walk]; by1 = g["py"][ids][:walk] + g["hy"][ids][:walk]
        boxhit = (bx1[:, None] >= X[None]) & (bx0[:, None] <= X[None]) & (by1[:, None] >= Y[None]) & (by0[:, None] <= Y[None])
        # (pixel-level box membership; a block is a candidate if any of its pixel centres' cells intersects:
        #  use block extents instead)
        per = {}
        for name, b in blk.items():
            nb = int(b.max()) + 1
            h = np.zeros((walk, nb), dtype=bool); c = np.zeros((walk, nb), dtype=bool)
            for q in range(nb):
                pix = b == q
                h[:, q] = contrib[:, pix].any(1)
                x_lo, x_hi, y_lo, y_hi = X[pix].min(), X[pix].max(), Y[pix].min(), Y[pix].max()
                c[:, q] = (bx1 >= x_lo) & (bx0 <= x_hi) & (by1 >= y_lo) & (by0 <= y_hi)
            c |= h
            hits[name] += int(h.sum()); cands[name] += int(c.sum())
            per[name] = (h, c)
        h88, _ = per["8x8"]
        cnt = contrib.reshape(walk, 16, 16)
        for q in range(4):
            sub = cnt[:, (q // 2) * 8:(q // 2) * 8 + 8, (q % 2) * 8:(q % 2) * 8 + 8].reshape(walk, 64).sum(1)
            hist_valid += np.bincount(sub[h88[:, q]], minlength=65)[:65]
        # two half warps walking their own lists side by side
        for big, small, nbig, key in (("8x8", "8x4", 4, "8x8|2x(8x4) halves"), ("8x4", "4x4", 8, "8x4|2x(4x4) halves")):
            hs, cs = per[small]
            for q in range(nbig):
                if big == "8x8":
                    lo, hi = (q // 2) * 4 + (q % 2), (q // 2) * 4 + 2 + (q % 2)   # 8x4 blocks: rows 2*(q//2), 2*(q//2)+1
                else:
                    lo, hi = (q // 2) * 4 + (q % 2) * 2, (q // 2) * 4 + (q % 2) * 2 + 1
                l_lo, l_hi = np.nonzero(cs[:, lo])[0], np.nonzero(cs[:, hi])[0]
                n = max(len(l_lo), len(l_hi))
                f_lo = np.zeros(n, dtype=bool); f_hi = np.zeros(n, dtype=bool)
                f_lo[:len(l_lo)] = hs[l_lo, lo]; f_hi[:len(l_hi)] = hs[l_hi, hi]
                paired[key + ": full iterations"] += int((f_lo | f_hi).sum())
                paired[key + ": all iterations"] += n
        # scalar kernel, one pixel per lane: warp = 8x4 block, four quarter warps = 4x2 blocks
        hs2, cs2 = per["4x2"]
        for wq in range(8):   # 8x4 warp blocks: column wq % 2, row wq // 2 (4 rows each)
            wx, wy = wq % 2, wq // 2
            subs = [(wy * 2 + dr) * 4 + (wx * 2 + dc) for dr in range(2) for dc in range(2)]
            lists = [np.nonzero(cs2[:, b])[0] for b in subs]
            n = max(len(l) for l in lists)
            full = np.zeros(n, dtype=bool)
            for l, b in zip(lists, subs):
                full[:len(l)] |= hs2[l, b]
            paired["8x4|4x(4x2) quarters (1 px / lane): full iterations"] += int(full.sum())
            paired["8x4|4x(4x2) quarters (1 px / lane): all iterations"] += n
        # four quarter warps (4x4 blocks) of an 8x8 warp block walking their own lists
        hs, cs = per["4x4"]
        first_pix = np.argmax(contrib, axis=0)   # front-most contributor per pixel (0 if none)
        has = contrib.any(0)
        for q in range(4):
            r0, c0 = (q // 2) * 2, (q % 2) * 2
            subs = [(r0 + dr) * 4 + (c0 + dc) for dr in range(2) for dc in range(2)]
            lists = [np.nonzero(cs[:, b])[0] for b in subs]
            n = max(len(l) for l in lists)
            full = np.zeros(n, dtype=bool)
            for l, b in zip(lists, subs):
                full[:len(l)] |= hs[l, b]
            paired["8x8|4x(4x4) quarters: full iterations"] += int(full.sum())
            paired["8x8|4x(4x4) quarters: all iterations"] += n
            pix = blk["8x8"] == q
            fe = np.unique(first_pix[pix & has])
            paired["8x8 hits containing a first contributor"] += len(fe)
        # a warp with FOUR pixels per lane: 16x8 (or 8x16) pixels, split into eight 4x4 blocks (4 lanes each) or four
        # 8x4 blocks (8 lanes each), every block walking its own list
        def lockstep(key, small, groups):
            hs_, cs_ = per[small]
            for subs in groups:
                lists = [np.nonzero(cs_[:, b])[0] for b in subs]
                n = max(len(l) for l in lists)
                full = np.zeros(n, dtype=bool)
                for l, b in zip(lists, subs):
                    full[:len(l)] |= hs_[l, b]
                paired[key + ": full iterations"] += int(full.sum())
                paired[key + ": all iterations"] += n
        # 4x4 block index = row4 * 4 + col4; 16x8 warp block = rows {2h, 2h+1} x all four columns
        lockstep("16x8|8x(4x4) eighths (4 px / lane)", "4x4", [[(2 * h + r) * 4 + c for r in range(2) for c in range(4)] for h in range(2)])
        # 8x16 warp block = all four rows x columns {2h, 2h+1}
        lockstep("8x16|8x(4x4) eighths (4 px / lane)", "4x4", [[r * 4 + 2 * h + c for r in range(4) for c in range(2)] for h in range(2)])
        # 8x4 block index = row4 * 2 + col8; 16x8 warp block = rows {2h, 2h+1} x both columns
        lockstep("16x8|4x(8x4) quarters (4 px / lane)", "8x4", [[(2 * h + r) * 2 + c for r in range(2) for c in range(2)] for h in range(2)])
    scale = (gx * gy) / float(len(tiles))
    print("sampled %d tiles (every %d) in %.1f s; numbers below are scaled to the whole frame" % (len(tiles), a.every, time.time() - t0))
    print("blended pairs NG = %.1f M, entries = %.2f M, walked entries = %.2f M" % (tot["pairs"] * scale / 1e6, tot["entries"] * scale / 1e6, tot["walk"] * scale / 1e6))
    for name in blk:
        npix = {"8x8": 64, "8x4": 32, "4x4": 16, "4x2": 8, "16x4": 64, "16x8": 128, "16x16": 256}[name]
        print("block %-6s hits %.2f M  candidates (bbox) %.2f M  valid pixels per hit %.1f of %d (%.0f %%)" % (
            name, hits[name] * scale / 1e6, cands[name] * scale / 1e6, tot["pairs"] / max(1, hits[name]), npix,
            100.0 * tot["pairs"] / max(1, hits[name]) / npix))
    for k, v in paired.items():
        print("%-60s %.2f M" % (k, v * scale / 1e6))
    cum = np.cumsum(hist_valid) / max(1, hist_valid.sum())
    print("8x8 hits by number of valid pixels: <=4: %.0f %%, <=8: %.0f %%, <=16: %.0f %%, <=32: %.0f %%" % (
        100 * cum[4], 100 * cum[8], 100 * cum[16], 100 * cum[32]))


if __name__ == "__main__":
    main()
