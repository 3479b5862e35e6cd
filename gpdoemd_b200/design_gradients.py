"""Gradients of the design criteria with respect to the design x.

The reference only evaluates a fixed candidate grid and takes ``np.argmax`` (demos/case_study.py:279-293).  To refine
that argmax continuously a caller needs d(criterion)/dx.  The expensive part -- the GP marginal (mu, S) of every rival
model and its x-derivative -- runs on the device (``taylor_first_order_dx`` -> ``gpd_marginal_dx``: the derivative
right-hand sides go through the same FP64 tensor-core triangular kernel as the prediction).  What is left is the chain
rule through the closed-form criteria of ``design_criteria.py:49-144`` for the handful of points being refined
(K x M^2 x E^3 x D flops): that is caller-side logic and is written here in NumPy, like the search loop itself.

    f, g = criterion_gradient('BH', mu, S, dmu, dS, noise_var, pps)      # f (K,), g (K, D)

with mu (K,M,E), S (K,M,E,E), dmu (K,M,E,D), dS (K,M,E,E,D).  All five criteria are covered (first-order marginals;
second-order refinement uses the derivative-free search)."""
import numpy as np

from .design_criteria import _reshape

SUPPORTED = ('HR', 'BH', 'BF', 'AW', 'JR')


def _quad_and_grad(iS, dS, r, dr):
    """q = r^T iS r and dq/dx_d = 2 r^T iS dr_d - r^T iS dS_d iS r  for S-inverse iS (K,E,E)."""
    w = np.einsum('kab,kb->ka', iS, r)
    q = np.einsum('ka,ka->k', r, w)
    dq = 2.0 * np.einsum('ka,kad->kd', w, dr) - np.einsum('ka,kabd,kb->kd', w, dS, w)
    return q, dq


def _jensen_renyi(mu, Sn, dmu, dS, pps, E):
    """JR (design_criteria.py:147-194) and its x-gradient.  With P_i = S_i^-1, l_i = log|S_i|:
        T1 = sum_i pi_i (E/2 log 4 pi + l_i / 2),   T2 = sum_i pi_i^2 2^(-E/2) |S_i|^(-1/2) + 2 sum_{i<j} pi_i pi_j e^(-phi_ij/2),
        phi_ij = mu_i'P_i mu_i + mu_j'P_j mu_j - m'(P_i+P_j)^-1 m + l_i + l_j + log|P_i+P_j|,   m = P_i mu_i + P_j mu_j,
        JR = E/2 log 2 pi - log T2 - T1;   dP = -P dS P,  dl = tr(P dS),  d log|R| = tr(R^-1 dR),  dR^-1 = -R^-1 dR R^-1."""
    K, M = mu.shape[:2]
    P = np.linalg.inv(Sn)
    ld = np.log(np.linalg.det(Sn))                                            # (K, M)
    dl = np.einsum('kmab,kmbad->kmd', P, dS)
    dP = -np.einsum('kmab,kmbcd,kmce->kmaed', P, dS, P)                       # (K, M, E, E, D)
    Pm = np.einsum('kmab,kmb->kma', P, mu)
    a = np.einsum('kma,kma->km', mu, Pm)
    da = 2.0 * np.einsum('kma,kmad->kmd', Pm, dmu) - np.einsum('kma,kmabd,kmb->kmd', Pm, dS, Pm)
    dPm = np.einsum('kmabd,kmb->kmad', dP, mu) + np.einsum('kmab,kmbd->kmad', P, dmu)
    T1 = np.sum(pps * (0.5 * E * np.log(4.0 * np.pi) + 0.5 * ld), axis=1)
    dT1 = 0.5 * np.einsum('m,kmd->kd', pps, dl)
    root = np.exp(-0.5 * ld) * 2.0 ** (-0.5 * E)                              # 2^(-E/2) |S_i|^(-1/2)
    T2 = np.sum(pps * pps * root, axis=1)
    dT2 = np.einsum('m,km,kmd->kd', pps * pps, -0.5 * root, dl)
    for i in range(M - 1):
        for j in range(i + 1, M):
            R, dR = P[:, i] + P[:, j], dP[:, i] + dP[:, j]
            Q = np.linalg.inv(R)
            m, dm = Pm[:, i] + Pm[:, j], dPm[:, i] + dPm[:, j]
            u = np.einsum('kab,kb->ka', Q, m)
            b = np.einsum('ka,ka->k', m, u)
            db = 2.0 * np.einsum('ka,kad->kd', u, dm) - np.einsum('ka,kabd,kb->kd', u, dR, u)
            c = np.log(np.linalg.det(R))
            dc = np.einsum('kab,kbad->kd', Q, dR)
            phi = a[:, i] + a[:, j] - b + ld[:, i] + ld[:, j] + c
            dphi = da[:, i] + da[:, j] - db + dl[:, i] + dl[:, j] + dc
            e = 2.0 * pps[i] * pps[j] * np.exp(-0.5 * phi)
            T2 += e
            dT2 += e[:, None] * (-0.5 * dphi)
    return 0.5 * E * np.log(2.0 * np.pi) - np.log(T2) - T1, -dT2 / T2[:, None] - dT1


def criterion_gradient(kind, mu, S, dmu, dS, noise_var=None, pps=None):
    assert kind in SUPPORTED, 'analytic design gradients cover %s' % (SUPPORTED,)
    mu, S, noise, pps, K, M, E = _reshape(np.asarray(mu, dtype=float), np.asarray(S, dtype=float), noise_var, pps)
    D = dmu.shape[-1]
    dmu = np.asarray(dmu, dtype=float).reshape(K, M, E, D)
    dS = np.asarray(dS, dtype=float).reshape(K, M, E, E, D)
    f = np.zeros(K)
    g = np.zeros((K, D))
    if kind == 'HR':                                      # design_criteria.py:49-63
        for i in range(M - 1):
            for j in range(i + 1, M):
                r, dr = mu[:, i] - mu[:, j], dmu[:, i] - dmu[:, j]
                f += np.einsum('ka,ka->k', r, r)
                g += 2.0 * np.einsum('ka,kad->kd', r, dr)
        return f, g
    Sn = S + noise                                        # :81, :114, :134 (the inputs are not modified)
    if kind in ('BH', 'AW'):
        iS = np.linalg.inv(Sn)
    if kind == 'BH':                                      # :66-92
        for i in range(M - 1):
            for j in range(i + 1, M):
                r, dr = mu[:, i] - mu[:, j], dmu[:, i] - dmu[:, j]
                t1 = np.einsum('kab,kba->k', Sn[:, i], iS[:, j]) + np.einsum('kab,kba->k', Sn[:, j], iS[:, i]) - 2.0 * E
                # d tr(S_i iS_j) = tr(dS_i iS_j) - tr(iS_j S_i iS_j dS_j)
                A_ij = np.einsum('kab,kbc,kcd->kad', iS[:, j], Sn[:, i], iS[:, j])
                A_ji = np.einsum('kab,kbc,kcd->kad', iS[:, i], Sn[:, j], iS[:, i])
                dt1 = (np.einsum('kabd,kba->kd', dS[:, i], iS[:, j]) - np.einsum('kab,kbad->kd', A_ij, dS[:, j]) +
                       np.einsum('kabd,kba->kd', dS[:, j], iS[:, i]) - np.einsum('kab,kbad->kd', A_ji, dS[:, i]))
                qi, dqi = _quad_and_grad(iS[:, i], dS[:, i], r, dr)
                qj, dqj = _quad_and_grad(iS[:, j], dS[:, j], r, dr)
                w = pps[i] * pps[j]
                f += w * (t1 + qi + qj)
                g += w * (dt1 + dqi + dqj)
        return 0.5 * f, 0.5 * g
    if kind == 'BF':                                      # :95-123
        for i in range(M - 1):
            for j in range(i + 1, M):
                r, dr = mu[:, i] - mu[:, j], dmu[:, i] - dmu[:, j]
                iA = np.linalg.inv(Sn[:, i] + Sn[:, j])
                dA = dS[:, i] + dS[:, j]
                B = np.einsum('kab,bc,kcd->kad', iA, noise, iA)          # iA N iA
                q, dq = _quad_and_grad(iA, dA, r, dr)
                f += np.einsum('ab,kba->k', noise, iA) + q
                g += -np.einsum('kab,kbad->kd', B, dA) + dq
        return f, g
    if kind == 'JR':
        return _jensen_renyi(mu, Sn, dmu, dS, pps, E)
    # AW  :126-144
    for i in range(M):
        w = np.zeros(K)
        dw = np.zeros((K, D))
        for j in range(M):
            r, dr = mu[:, i] - mu[:, j], dmu[:, i] - dmu[:, j]
            q, dq = _quad_and_grad(iS[:, i], dS[:, i], r, dr)
            e = np.exp(-0.5 * q)
            w += e
            dw += e[:, None] * (-0.5 * dq)
        f += pps[i] / w
        g += -pps[i] * dw / (w ** 2)[:, None]
    return f, g
