"""Explicit-index restatement (numpy float64, plain loops / index arithmetic) of the third-party primitives the hot
path leans on  --  TEST INFRASTRUCTURE ONLY (same rules as oracle/sqldepth_oracle.py).

The reference calls PyTorch for these steps; the CUDA kernels re-implement them from the index rules below, so the
rules themselves are pinned here against the library (tests/test_oracle_golden.py::test_explicit_*):

  upsample_bilinear      F.interpolate(x, [H, W], mode="bilinear", align_corners=False)     trainer.py:395-396
                         src = (dst + 0.5) * (in / out) - 0.5, clamped at 0; second tap clamped to the last row / column
                         (ATen UpSample.h: area_pixel_compute_source_index)
  grid_sample_border     F.grid_sample(src, grid, padding_mode="border", align_corners=True)  trainer.py:431-435
                         ix = (gx + 1) / 2 * (W - 1), clipped to [0, W - 1]; bilinear; taps beyond the frame have weight 0
                         (ATen GridSampler.h: grid_sampler_compute_source_index + clip_coordinates)
  ssim_reflect           layers.py:13-46: ReflectionPad2d(3) (i < 0 -> -i, i >= n -> 2 (n - 1) - i), 7x7 mean filters,
                         clamp((1 - n / d) / 2, 0, 1) with C1 = 0.01^2, C2 = 0.03^2
  pixel_softmax_summary  networks/layers.py:17-19: softmax over PIXELS of x^T K, summaries
  bin_centers            networks/depth_decoder_QTR.py:52-66 (norm == 'linear')
"""
import numpy as np


def upsample_bilinear(x, H, W):
    """x [h, w] -> [H, W]"""
    h, w = x.shape
    out = np.empty((H, W), dtype=np.float64)
    for v in range(H):
        sy = max((v + 0.5) * (h / H) - 0.5, 0.0)
        i0 = int(np.floor(sy))
        i1 = min(i0 + 1, h - 1)
        ly = sy - i0
        for u in range(W):
            sx = max((u + 0.5) * (w / W) - 0.5, 0.0)
            j0 = int(np.floor(sx))
            j1 = min(j0 + 1, w - 1)
            lx = sx - j0
            out[v, u] = (1 - ly) * ((1 - lx) * x[i0, j0] + lx * x[i0, j1]) + ly * ((1 - lx) * x[i1, j0] + lx * x[i1, j1])
    return out


def grid_sample_border(src, grid):
    """src [C, H, W], grid [H', W', 2] normalised (x, y) -> [C, H', W']"""
    C, H, W = src.shape
    Ho, Wo, _ = grid.shape
    out = np.empty((C, Ho, Wo), dtype=np.float64)
    for v in range(Ho):
        for u in range(Wo):
            ix = min(max((grid[v, u, 0] + 1) / 2 * (W - 1), 0.0), W - 1.0)
            iy = min(max((grid[v, u, 1] + 1) / 2 * (H - 1), 0.0), H - 1.0)
            x0, y0 = int(np.floor(ix)), int(np.floor(iy))
            fx, fy = ix - x0, iy - y0
            acc = np.zeros(C)
            for yy, wy in ((y0, 1 - fy), (y0 + 1, fy)):
                for xx, wx in ((x0, 1 - fx), (x0 + 1, fx)):
                    if 0 <= yy < H and 0 <= xx < W:          # out-of-frame taps only occur with weight 0
                        acc += wy * wx * src[:, yy, xx]
            out[:, v, u] = acc
    return out


def _reflect(i, n):
    if i < 0:
        return -i
    if i >= n:
        return 2 * (n - 1) - i
    return i


def ssim_reflect(x, y, radius=3):
    """x, y [H, W] -> SSIM loss map [H, W]"""
    H, W = x.shape
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    k = (2 * radius + 1) ** 2
    out = np.empty((H, W), dtype=np.float64)
    for v in range(H):
        rows = [_reflect(v + d, H) for d in range(-radius, radius + 1)]
        for u in range(W):
            cols = [_reflect(u + d, W) for d in range(-radius, radius + 1)]
            a = x[np.ix_(rows, cols)]
            b = y[np.ix_(rows, cols)]
            mx, my = a.sum() / k, b.sum() / k
            sx = (a * a).sum() / k - mx * mx
            sy = (b * b).sum() / k - my * my
            sxy = (a * b).sum() / k - mx * my
            n = (2 * mx * my + C1) * (2 * sxy + C2)
            d = (mx * mx + my * my + C1) * (sx + sy + C2)
            out[v, u] = min(max((1 - n / d) / 2, 0.0), 1.0)
    return out


def pixel_softmax_summary(x, K):
    """x [E, n], K [Q, E] -> (energy [Q, n], summary [Q, E]); the softmax runs over the n pixels of each query"""
    E, n = x.shape
    Q = K.shape[0]
    energy = np.empty((Q, n))
    summary = np.empty((Q, E))
    for q in range(Q):
        for p in range(n):
            energy[q, p] = sum(x[e, p] * K[q, e] for e in range(E))
        m = energy[q].max()
        w = np.exp(energy[q] - m)
        w /= w.sum()
        for e in range(E):
            summary[q, e] = sum(w[p] * x[e, p] for p in range(n))
    return energy, summary


def bin_centers(r, min_val, max_val):
    """r [D] regressor output -> bin centres [D]"""
    y = np.maximum(r, 0.0) + 0.1
    y = y / y.sum()
    edge = min_val
    out = np.empty_like(y)
    for d in range(len(y)):
        nxt = edge + (max_val - min_val) * y[d]
        out[d] = 0.5 * (edge + nxt)
        edge = nxt
    return out
