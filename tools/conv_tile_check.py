"""Host-side model of the register-tiled Conv2D kernels for single-channel images (csrc/conv.cu: k_conv1_fwd_tiled,
k_conv1_bwd_tiled).  It restates, in NumPy and pure Python, exactly what the kernels index:
  * ct_stride / ct_geom: row pitches that are multiples of 4 floats with an odd quarter;
  * the bordered, channel-planar gradient tile gb[o][y + K-1][x + 4] and the padded image rows the backward kernel stages;
  * dx from 4 x 4 register tiles walking the (4 + K-1) x 8 window above the tile (filter row i = r + K-1 - wr, tap j = c + 4 - wc);
  * dw / db with one (channel, filter row) pair per warp and one output row per lane, sliding four positions at a time;
  * the forward's four-positions-per-thread strips;
and checks (a) the results against the oracle's literal lowering (im2col + GEMM + col2im), (b) that no access leaves its buffer,
(c) that the 16-byte shared-memory accesses of a quarter warp fall into 8 distinct bank groups wherever the kernel's comments say so.
Pure CPU: used by tests/test_tile_layouts.py."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle_np as O  # noqa: E402

COL_BORDER = 4


def ct_stride(need):
    s = (need + 3) & ~3
    return s if (s >> 2) & 1 else s + 4


def ct_geom(H, W, K, co):
    gs = ct_stride(((W + 3) // 4) * 4 + COL_BORDER)
    gr = ((H + 3) // 4) * 4 + (K - 1)
    gcs = gr * gs + 4
    is_ = ct_stride(W + 4)
    return dict(gs=gs, gr=gr, gcs=gcs, is_=is_, img_sz=(H + 1) * is_)


def stage(im_b, g_b, K, co, ge):
    """one sample: image [H, W] -> padded rows; gradient [P, co] (NHWC) -> planar tile with zero border; flat float arrays"""
    H, W = im_b.shape
    Ho, Wo = H - K + 1, W - K + 1
    img = np.zeros(ge["img_sz"], np.float32)
    for r in range(H):
        img[r * ge["is_"]: r * ge["is_"] + W] = im_b[r]
    gb = np.zeros(co * ge["gcs"], np.float32)
    g3 = g_b.reshape(Ho, Wo, co)
    for o in range(co):
        for y in range(Ho):
            base = o * ge["gcs"] + (y + K - 1) * ge["gs"] + COL_BORDER
            gb[base: base + Wo] = g3[y, :, o]
    return img, gb


def bwd_model(im, filt, g, K, co):
    B, _, H, W = im.shape
    Ho, Wo = H - K + 1, W - K + 1
    ge = ct_geom(H, W, K, co)
    dx = np.zeros((B, 1, H, W), np.float32)
    dw = np.zeros((K, K, co), np.float64)
    db = np.zeros(co, np.float64)
    touched = []
    for b in range(B):
        img, gb = stage(im[b, 0], g[b].reshape(Ho * Wo, co), K, co, ge)
        # dx: thread (tu, tv) owns pixels [4tu, 4tu+4) x [4tv, 4tv+4)
        for tu in range((H + 3) // 4):
            for tv in range((W + 3) // 4):
                acc = np.zeros((4, 4), np.float32)
                for o in range(co):
                    for wr in range(4 + K - 1):
                        a0 = o * ge["gcs"] + (4 * tu + wr) * ge["gs"] + 4 * tv
                        touched.append(("gb", a0 + 7, gb.size))
                        w = gb[a0: a0 + 8]
                        for r in range(4):
                            i = r + (K - 1) - wr
                            if not 0 <= i < K:
                                continue
                            for c in range(4):
                                for j in range(K):
                                    acc[r, c] += w[c + COL_BORDER - j] * filt[i, j, o]
                for r in range(4):
                    for c in range(4):
                        if 4 * tu + r < H and 4 * tv + c < W:
                            dx[b, 0, 4 * tu + r, 4 * tv + c] = acc[r, c]
        # dw / db: lane = output row, pair (o, i); slide four positions at a time
        for o in range(co):
            for i in range(K):
                for lane in range(Ho):
                    ip = (lane + i) * ge["is_"]
                    gp = o * ge["gcs"] + (lane + K - 1) * ge["gs"] + COL_BORDER
                    for x in range(0, Wo, 4):
                        touched.append(("img", ip + x + 7, img.size))
                        touched.append(("gb", gp + x + 3, gb.size))
                        w, gq = img[ip + x: ip + x + 8], gb[gp + x: gp + x + 4]
                        for xx in range(4):
                            for j in range(K):
                                dw[i, j, o] += float(w[xx + j]) * float(gq[xx])
                        if i == 0:
                            db[o] += float(gq.sum())
    return dx, dw.astype(np.float32), db.astype(np.float32), touched, ge


def fwd_model(im, filt, bias, K, co):
    B, _, H, W = im.shape
    Ho, Wo = H - K + 1, W - K + 1
    out = np.zeros((B, Ho * Wo, co), np.float32)
    raw = np.concatenate([im.reshape(-1), np.zeros(8, np.float32)])
    touched = []
    for b in range(B):
        for y in range(Ho):
            for x0 in range(0, Wo, 4):
                acc = np.zeros((4, co), np.float32)
                for i in range(K):
                    a0 = b * H * W + (y + i) * W + x0
                    touched.append(("raw", a0 + 7, raw.size))
                    w = raw[a0: a0 + 8]
                    for j in range(K):
                        for q in range(4):
                            acc[q] += w[q + j] * filt[i, j]
                for q in range(4):
                    if x0 + q < Wo:
                        out[b, y * Wo + x0 + q] = acc[q] + bias
    return out, touched


def bank_groups(float_addresses):
    """16-byte accesses of a quarter warp: the bank group (of 8) each address falls into"""
    return [(a // 4) % 8 for a in float_addresses]


def check():
    bad = []
    rng = np.random.default_rng(5)
    for (B, H, W, K, co) in ((2, 12, 12, 5, 6), (1, 9, 8, 3, 8), (2, 7, 16, 5, 3), (1, 28, 28, 5, 6)):
        im = rng.standard_normal((B, 1, H, W)).astype(np.float32)
        filt = rng.standard_normal((K, K, co)).astype(np.float32)
        bias = rng.standard_normal(co).astype(np.float32)
        want_out, col2d = O.conv2d_fwd(im, filt, bias, 1, K, 1, 0)
        g = rng.standard_normal(want_out.shape).astype(np.float32)
        wdx, wdw, wdb = O.conv2d_bwd(im, filt, col2d, g, 1, K, 1, 0, literal=False)
        if (B, H) == (1, 28):  # the LeNet geometry: pitches and bank groups only (the pure-Python model is slow at this size)
            ge = ct_geom(H, W, K, co)
            if (ge["gs"], ge["gr"], ge["is_"]) != (36, 32, 36):
                bad.append(f"LeNet geometry {ge}")
            for name, pitch in (("gb rows (dw: lanes differ in y)", ge["gs"]), ("image rows (dw)", ge["is_"])):
                if sorted(bank_groups([y * pitch for y in range(8)])) != list(range(8)):
                    bad.append(f"{name}: pitch {pitch} is not conflict free")
            if sorted(bank_groups([4 * tv for tv in range(8)])) != list(range(8)):
                bad.append("dx window loads of 8 neighbouring tiles collide")
            continue
        dx, dw, db, touched, ge = bwd_model(im, filt, g.reshape(B, -1, co), K, co)
        out, touched_f = fwd_model(im, filt, bias, K, co)
        for what, got, want in (("dx", dx, wdx), ("dw", dw, wdw), ("db", db, wdb), ("out", out.reshape(want_out.shape), want_out)):
            if got.shape != want.shape or O.rel_err(got, want) > 2e-5:
                bad.append(f"{what} differs from the oracle at B={B} H={H} W={W} K={K} co={co}: {O.rel_err(got, want) if got.shape == want.shape else got.shape}")
        for name, last, size in touched + touched_f:
            if last >= size:
                bad.append(f"{name}: access up to {last} of {size} at H={H} W={W} K={K}")
                break
        for pitch in (ge["gs"], ge["is_"]):
            if pitch % 4 or not (pitch >> 2) & 1:
                bad.append(f"pitch {pitch}: not a multiple of 4 with an odd quarter")
    return bad


if __name__ == "__main__":
    problems = check()
    print("\n".join(problems) if problems else "conv tile model: all checks passed")
    sys.exit(1 if problems else 0)
