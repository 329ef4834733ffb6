"""Explicit restatements of the third-party numeric primitives on the hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
baseline legs may use it (as the checker, never as the thing shipped).

The reference (/root/reference, pure Python) delegates its arithmetic to numpy /
scipy calls whose source is not under /root/reference:

  * ``scipy.ndimage.gaussian_filter``   Utils/ScanMatcher_OGBased.py:42
  * ``np.sum(axis=2)`` / ``np.sum``      Utils/ScanMatcher_OGBased.py:130,137,141
  * ``np.unique(axis=0)``               Utils/ScanMatcher_OGBased.py:120
  * ``np.random.choice(..., p=...)``    Utils/ScanMatcher_OGBased.py:138, Algorithm/FastSlam.py:59
  * ``np.linspace``                     Utils/ScanMatcher_OGBased.py:82

The CUDA kernels cannot call numpy, so each primitive's exact operation order is
restated here in scalar terms (this is what the kernels implement) and pinned
against the real numpy/scipy call in ``tests/test_oracle_primitives.py``.
Versions in this image: numpy 2.3.5, scipy 1.18.1 (the reference pins none).
"""
import numpy as np

PW_BLOCK = 128  # numpy's PW_BLOCKSIZE


def pairwise_sum(a):
    """numpy's float64 add-reduction over a contiguous 1-D run (npy pairwise_sum).

    n < 8: sequential from 0.0; n <= 128: eight running lanes then a fixed tree,
    then the n % 8 tail added sequentially; n > 128: split at n/2 rounded down to a
    multiple of 8 and recurse.
    """
    a = np.asarray(a, dtype=np.float64)
    n = a.shape[0]
    if n < 8:
        res = np.float64(0.0)
        for i in range(n):
            res = res + a[i]
        return res
    if n <= PW_BLOCK:
        r = [a[i] for i in range(8)]
        m = n - (n % 8)
        for i in range(8, m, 8):
            for l in range(8):
                r[l] = r[l] + a[i + l]
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
        for i in range(m, n):
            res = res + a[i]
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return pairwise_sum(a[:n2]) + pairwise_sum(a[n2:])


def pairwise_leaves(n, off=0, out=None):
    """Leaf runs (offset, length <= 128) of the pairwise recursion, in tree order."""
    if out is None:
        out = []
    if n <= PW_BLOCK:
        out.append((off, n))
        return out
    n2 = n // 2
    n2 -= n2 % 8
    pairwise_leaves(n2, off, out)
    pairwise_leaves(n - n2, off + n2, out)
    return out


def gaussian_weights(sigma, truncate=4.0):
    """scipy.ndimage gaussian kernel: radius int(truncate*sigma+0.5), normalised."""
    sd = float(sigma)
    radius = int(truncate * sd + 0.5)
    sigma2 = sigma * sigma
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / sigma2 * x ** 2)
    phi = phi / phi.sum()
    return phi[::-1].copy(), radius


def _reflect_index(i, n):
    """scipy mode='reflect' (half-sample symmetric): ... 1 0 | 0 1 2 ... n-1 | n-1 n-2 ..."""
    period = 2 * n
    i = np.mod(i, period)
    return np.where(i >= n, period - 1 - i, i)


def correlate1d_symmetric(x, w, radius, axis):
    """scipy NI_Correlate1D, symmetric-kernel branch, restated.

    out = x[c]*w[r]; for j = -r..-1: out = out + (x[c+j] + x[c-j]) * w[r+j]
    (pair add, multiply, accumulate; no fused multiply-add)."""
    x = np.moveaxis(np.asarray(x, dtype=np.float64), axis, -1)
    n = x.shape[-1]
    idx = np.arange(n)
    out = x * w[radius]
    for j in range(-radius, 0):
        lo = x[..., _reflect_index(idx + j, n)]
        hi = x[..., _reflect_index(idx - j, n)]
        out = out + (lo + hi) * w[radius + j]
    return np.moveaxis(out, -1, axis)


def gaussian_filter_reflect(x, sigma):
    """scipy.ndimage.gaussian_filter(x, sigma) for 2-D float64, default mode."""
    w, r = gaussian_weights(sigma)
    tmp = correlate1d_symmetric(x, w, r, axis=0)
    return correlate1d_symmetric(tmp, w, r, axis=1)


def unique_xy(xi, yi):
    """np.unique(column_stack((xi, yi)), axis=0): distinct pairs sorted by (x, then y)."""
    key = xi.astype(np.int64) * (1 << 20) + yi.astype(np.int64)
    key = np.unique(key)
    return key >> 20, key & ((1 << 20) - 1)


def legacy_choice_index(p, u):
    """RandomState.choice(arange(n), size, p=p) given the uniforms it would draw.

    cdf = p.cumsum() (sequential); cdf /= cdf[-1]; searchsorted(u, side='right')."""
    cdf = np.cumsum(np.asarray(p, dtype=np.float64))
    cdf = cdf / cdf[-1]
    return np.searchsorted(cdf, u, side='right')


def linspace(start, stop, num):
    """np.linspace(start, stop, num) for num > 1, endpoint=True (step != 0)."""
    step = (stop - start) / (num - 1)
    y = np.arange(0, num).astype(np.float64) * step + start
    y[-1] = stop
    return y
