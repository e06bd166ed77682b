"""Shared tolerance helpers of the GPU parity tests.  north_star: "within 1e-4 relative fp32" — ELEMENTWISE:
|got - want| <= atol + 1e-4 |want| for every element, with a small absolute floor (atol) for values near zero, where
"relative" has no meaning in fp32 (a sum of O(1) terms that cancels to 1e-6 carries 1e-7 of rounding noise whatever
the summation order)."""
import numpy as np

RTOL = 1e-4


def close(got, want, what, rtol=RTOL, atol=2e-6):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    err = np.abs(got - want)
    bad = err > atol + rtol * np.abs(want)
    worst = float((err / (np.abs(want) + atol / rtol)).max()) if err.size else 0.0
    print(f"    {what}: max elementwise rel err {worst:.3e}")
    assert not bad.any(), (f"{what}: {int(bad.sum())} of {got.size} elements outside {rtol:g} elementwise "
                           f"(worst abs {float(err.max()):.3e} at value {float(want.flat[int(err.argmax())]):.3e})")


def close_params(got, want, what, lr, rounds, rtol=RTOL, atol=2e-6, max_frac=5e-5):
    """Parameters after `rounds` AdamW steps: elementwise 1e-4 like `close`, plus an explicit, counted and bounded list
    of AdamW outliers.  AdamW moves an element by lr * m / (sqrt(v) + eps) whatever the gradient's magnitude, so an element
    whose gradient is zero to within fp32 summation noise can step the other way than in the reference: at most
    max(3, max_frac * size) such elements (measured: 3 of 73 484 actor weights after 5 unsynchronised SAC rounds), each
    within 2 lr per step; they are printed."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    err = np.abs(got - want)
    bad = np.flatnonzero(err > atol + rtol * np.abs(want))
    worst = float((err / (np.abs(want) + atol / rtol)).max())
    print(f"    {what}: max elementwise rel err {worst:.3e}; AdamW outliers (index, got, want): "
          f"{[(int(i), float(got.flat[i]), float(want.flat[i])) for i in bad[:8]]}{' ...' if bad.size > 8 else ''}")
    allowed = max(3, int(max_frac * got.size))
    assert bad.size <= allowed, f"{what}: {bad.size} of {got.size} elements outside {rtol:g} elementwise (allowed AdamW outliers: {allowed})"
    if bad.size:
        assert float(err.flat[bad].max()) <= 2.02 * lr * rounds + atol, \
            f"{what}: outlier off by {float(err.flat[bad].max()):.3e}, beyond the AdamW bound 2 lr rounds"
