"""Shared tolerance helper of the parity tests.

BASELINE.json north_star: projections, IoUs, losses and gradients within 1e-5 relative fp32
tolerance.  A CUDA fp32 result and a torch-CPU fp32 result are BOTH a few ulps-times-condition away
from the true value, so their mutual distance is not the yardstick; the yardstick is a float64
evaluation of the same formulation (the oracle restatement, dtype-generic, pinned to the reference
at fp32 by tests/test_oracle_*.py, or the reference's own source text where it runs in float64):

    max |got - f64|  <=  1e-5 * max |f64|

If the reference's own fp32 evaluation is farther than that from the float64 value (an
ill-conditioned case: e.g. a rotation gradient summed over thousands of signed per-point terms, where
the fp32 per-term arithmetic the contract prescribes is the error), the case is printed and the bound
becomes that distance plus the 1e-5 allowance (i.e. what "within 1e-5 of the fp32 reference" implies).
"""
import numpy as np
import torch

TOL = 1e-5


def np64(a):
    if torch.is_tensor(a):
        return a.detach().cpu().double().numpy()
    return np.asarray(a, dtype=np.float64)


def d64(*tensors):
    """float64 leaf copies of torch / numpy inputs (requires_grad preserved)."""
    out = []
    for t in tensors:
        if t is None:
            out.append(None)
            continue
        rg = torch.is_tensor(t) and t.requires_grad
        x = torch.as_tensor(np64(t)).clone()
        out.append(x.requires_grad_(True) if rg else x)
    return out if len(out) > 1 else out[0]


def close64(got, ref64, ref32=None, tol=TOL, what=''):
    g, r = np64(got), np64(ref64)
    assert g.shape == r.shape, (what, g.shape, r.shape)
    if r.size == 0:
        return True
    scale = max(float(np.abs(r).max()), 1e-30)
    err = float(np.abs(g - r).max())
    bound = tol * scale
    if ref32 is not None:
        e32 = float(np.abs(np64(ref32) - r).max())
        if e32 > bound:
            print(f'[parity] {what}: the fp32 reference is {e32 / scale:.2e} (relative) from its float64 value; '
                  f'bound raised from {tol:.0e} to that + {tol:.0e}')
            bound = e32 + tol * scale
    ok = err <= bound
    if not ok:
        i = np.unravel_index(np.abs(g - r).argmax(), r.shape)
        print(f'[parity] {what}: max |got - f64| = {err:.3e} = {err / scale:.2e} of scale {scale:.3e} at {i}: '
              f'got {g[i]!r} f64 {r[i]!r} (bound {bound:.3e})')
    return ok
