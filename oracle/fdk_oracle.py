"""fp64 restatement of FDK for circular cone-beam scans.  TEST INFRASTRUCTURE ONLY (see tsp_oracle.c).

The reference reaches FDK through ``astra.experimental.accumulate_FDK`` (``tomosipo/astra.py:374-406``); ASTRA is
absent, so -- as for the projectors -- parity is anchored on the published algorithm (Feldkamp, Davis, Kress 1984;
Kak & Slaney eq. 3.176 ff; Parker 1982 for short scans) and on closed forms (tests/test_fdk.py: a centred ball is
reconstructed to its density).  **Parity against ASTRA's own FDK output: unpinned** (no FDK output exists anywhere in
the reference; ``tests/test_astra.py:66-84`` only checks that the call runs).

This restatement shares no code with ``tomosipo_b200.algorithms.fdk``: the ramp filter is a direct spatial
convolution with the band-limited kernel (no FFT), in float64, and the backprojection is the C oracle's.

    f(x) = 1/2 * integral over beta of  SOD^2 / (SOD - depth(x, beta))^2  *  q_beta(U(x), V(x))  d beta
    q    = (cos-weighted projection) convolved along u with h / tau,   tau = detector pitch scaled to the isocentre
    h[0] = 1/4,  h[n odd] = -1 / (pi n)^2,  h[n even] = 0

The C oracle's backprojector computes  V_vox * SDD^2 / (|u||v| (SOD - depth)^2) * interp(.)  per angle, so the
remaining factor per angle is  redundancy * d_beta * SOD^2 |u||v| / (SDD^2 V_vox tau)  with redundancy = 1/2 for a
full circle (every line is measured twice) and Parker's weight w(beta, gamma) for a short scan.
"""
import numpy as np

from . import oracle as O


def circular_table(vectors, vol_centre_xyz):
    """Per-angle scan constants recovered from ASTRA cone_vec rows (x, y, z order)."""
    v = np.asarray(vectors, dtype=np.float64)
    src, det, du, dv = v[:, 0:3], v[:, 3:6], v[:, 6:9], v[:, 9:12]
    pu, pv = np.linalg.norm(du, axis=1), np.linalg.norm(dv, axis=1)
    nrm = np.cross(du, dv)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    sdd = np.abs(np.sum((det - src) * nrm, axis=1))
    sod = np.abs(np.sum((np.asarray(vol_centre_xyz, dtype=np.float64) - src) * nrm, axis=1))
    foot = src + np.sum((det - src) * nrm, axis=1, keepdims=True) * nrm
    ppu = np.sum((foot - det) * du, axis=1) / pu ** 2
    ppv = np.sum((foot - det) * dv, axis=1) / pv ** 2
    beta = np.unwrap(np.arctan2(src[:, 0], -src[:, 1]))       # ASTRA: src = SOD (sin beta, -cos beta, 0)
    return dict(pu=pu, pv=pv, sdd=sdd, sod=sod, ppu=ppu, ppv=ppv, beta=beta)


def parker_weights(beta, gamma):
    """Parker (1982) redundancy weights w[angle, column] for source angles ``beta`` (ascending, uniform) covering
    pi + 2 delta, and fan angles ``gamma`` [column] in Parker's sign convention (the conjugate of (beta, gamma) is
    (beta + pi + 2 gamma, -gamma)).  delta is taken from the scan range, so any over-scan up to a full circle is used."""
    beta = np.asarray(beta, dtype=np.float64)
    b = beta - beta[0]
    step = float(np.mean(np.diff(beta)))
    rng = b[-1] + step                                         # the sampled arc, one step per angle
    delta = 0.5 * (rng - np.pi)
    g = np.asarray(gamma, dtype=np.float64)[None, :]
    if delta < np.abs(g).max() - 1e-9:
        raise ValueError(f"scan range {rng:.4f} rad is shorter than pi + the fan angle {2 * np.abs(g).max():.4f}")
    b = b[:, None]
    w = np.ones((len(beta), g.shape[1]))
    with np.errstate(divide="ignore", invalid="ignore"):
        ramp_up = np.sin(0.25 * np.pi * b / (delta - g)) ** 2
        ramp_dn = np.sin(0.25 * np.pi * (np.pi + 2 * delta - b) / (delta + g)) ** 2
    up = b < 2 * delta - 2 * g
    dn = b > np.pi - 2 * g
    w = np.where(up, ramp_up, w)
    w = np.where(dn, ramp_dn, w)
    return np.clip(np.nan_to_num(w, nan=0.0), 0.0, 1.0)


def fdk(kind_vectors, vol_shape_zyx, win_min_xyz, win_max_xyz, det_shape_vu, y, short_scan=False):
    """FDK reconstruction (float64) of projections ``y`` [V, A, U] for ASTRA cone_vec rows ``kind_vectors``."""
    vec = np.asarray(kind_vectors, dtype=np.float64)
    V, U = det_shape_vu
    A = vec.shape[0]
    centre = 0.5 * (np.asarray(win_min_xyz, dtype=np.float64) + np.asarray(win_max_xyz, dtype=np.float64))
    t = circular_table(vec, centre)
    vox = float(np.prod((np.asarray(win_max_xyz, dtype=np.float64) - np.asarray(win_min_xyz, dtype=np.float64))
                        / np.asarray(vol_shape_zyx[::-1], dtype=np.float64)))
    y = np.asarray(y, dtype=np.float64)
    iu = np.arange(U) + 0.5 - U / 2
    iv = np.arange(V) + 0.5 - V / 2
    n = np.arange(-(U - 1), U)
    h = np.where(n == 0, 0.25, np.where(n % 2 != 0, -1.0 / (np.pi * np.where(n == 0, 1, n)) ** 2, 0.0))
    step = np.abs(np.mean(np.diff(t["beta"]))) if A > 1 else 2 * np.pi
    q = np.zeros_like(y)
    for a in range(A):
        up = (iu - t["ppu"][a]) * t["pu"][a]
        vp = (iv - t["ppv"][a]) * t["pv"][a]
        cosw = t["sdd"][a] / np.sqrt(t["sdd"][a] ** 2 + up[None, :] ** 2 + vp[:, None] ** 2)
        p1 = y[:, a, :] * cosw
        if short_scan:
            red = parker_weights(t["beta"], -np.arctan2(up, t["sdd"][a]))[a][None, :]
        else:
            red = 0.5
        p1 = p1 * red
        tau = t["pu"][a] * t["sod"][a] / t["sdd"][a]
        for r in range(V):
            q[r, a, :] = np.convolve(p1[r], h)[U - 1: 2 * U - 1] / tau
        q[:, a, :] *= step * t["sod"][a] ** 2 * t["pu"][a] * t["pv"][a] / (t["sdd"][a] ** 2 * vox)
    P = O.OracleProjector(O.CONE_VEC, vol_shape_zyx, win_min_xyz, win_max_xyz, det_shape_vu, vec)
    return P.bp(q)
