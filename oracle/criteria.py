"""Restatement of GPdoemd/design_criteria.py:49-238 (TEST INFRASTRUCTURE ONLY).

Differences from the reference, on purpose: the oracle never mutates its inputs (the reference's
BH/BF do ``s2 += noise_var`` on a reshape view, design_criteria.py:81,114 -- SURVEY quirk 2); the
numbers returned for a given (mu, s2, noise_var, pps) are the same.
"""
import numpy as np


def _reshape(mu, s2, noise_var=None, pps=None):          # design_criteria.py:202-238
    if mu.ndim == 2:
        mu = np.expand_dims(mu, axis=2)
    n, M, E = mu.shape
    if noise_var is None:
        noise_var = np.zeros((E, E))
    if isinstance(noise_var, (int, float)):
        noise_var = np.array([noise_var])
    if noise_var.ndim == 1:
        tmp = np.eye(E)
        np.fill_diagonal(tmp, noise_var)
        noise_var = tmp
    assert noise_var.shape == (E, E)
    assert np.all(np.diag(noise_var) >= 0.)
    s2 = s2.reshape((n, M, E, E))
    if pps is None:
        pps = np.ones(M)
    if isinstance(pps, (list, tuple)):
        pps = np.array(pps)
    assert pps.shape == (M,) and np.all(pps >= 0)
    pps = pps / np.sum(pps)
    return mu, s2, noise_var, pps, n, M, E


def HR(mu, s2, noise_var=None, pps=None):                # :49-63
    mu, _, _, _, n, M, _ = _reshape(mu, s2, None, None)
    dc = np.zeros(n)
    for i in range(M - 1):
        for j in range(i + 1, M):
            dc += np.sum((mu[:, i] - mu[:, j]) ** 2, axis=1)
    return dc


def BH(mu, s2, noise_var=None, pps=None):                # :66-92
    mu, s2, noise_var, pps, n, M, E = _reshape(mu, s2, noise_var, pps)
    S = s2 + noise_var
    iS = np.linalg.inv(S)
    dc = np.zeros(n)
    for i in range(M - 1):
        for j in range(i + 1, M):
            t1 = np.trace(np.matmul(S[:, i], iS[:, j]) + np.matmul(S[:, j], iS[:, i])
                          - 2 * np.eye(E), axis1=1, axis2=2)
            r1 = np.expand_dims(mu[:, i] - mu[:, j], 2)
            t2 = np.sum(r1 * np.matmul(iS[:, i] + iS[:, j], r1), axis=(1, 2))
            dc += pps[i] * pps[j] * (t1 + t2)
    return 0.5 * dc


def BF(mu, s2, noise_var=None, pps=None):                # :95-123
    mu, s2, noise_var, _, n, M, _ = _reshape(mu, s2, noise_var, None)
    S = s2 + noise_var
    dc = np.zeros(n)
    for i in range(M - 1):
        for j in range(i + 1, M):
            iSij = np.linalg.inv(S[:, i] + S[:, j])
            t1 = np.trace(np.matmul(noise_var, iSij), axis1=1, axis2=2)
            r1 = np.expand_dims(mu[:, i] - mu[:, j], 2)
            t2 = np.sum(r1 * np.matmul(iSij, r1), axis=(1, 2))
            dc += t1 + t2
    return dc


def AW(mu, s2, noise_var=None, pps=None):                # :126-144
    mu, s2, noise_var, pps, n, M, _ = _reshape(mu, s2, noise_var, pps)
    iS = np.linalg.inv(s2 + noise_var)
    dc = np.zeros((n, M))
    for i in range(M):
        for j in range(M):
            r1 = np.expand_dims(mu[:, i] - mu[:, j], 2)
            t1 = np.sum(r1 * np.matmul(iS[:, i], r1), axis=(1, 2))
            dc[:, i] += np.exp(-0.5 * t1)
    return np.sum((1. / dc) * pps, axis=1)


def JR(mu, s2, noise_var=None, pps=None):                # :147-194
    mu, s2, noise_var, pps, _, M, E = _reshape(mu, s2, noise_var, pps)
    S = s2 + noise_var
    iS = np.linalg.inv(S)
    dS = np.linalg.det(S)
    ldS = np.log(dS)
    T1 = np.sum(pps * ((E / 2) * np.log(4 * np.pi) + 0.5 * ldS), axis=1)
    T2 = np.sum(pps * pps / (2 ** (E / 2.) * np.sqrt(dS)), axis=1)
    for i in range(M):
        mi = np.expand_dims(mu[:, i], 2)
        iSmi = np.matmul(iS[:, i], mi)
        miiSmi = np.sum(mi * iSmi, axis=(1, 2))
        for j in range(i + 1, M):
            mj = np.expand_dims(mu[:, j], 2)
            iSmj = np.matmul(iS[:, j], mj)
            mjiSmj = np.sum(mj * iSmj, axis=(1, 2))
            iSiS = iS[:, i] + iS[:, j]
            iiSiS = np.linalg.inv(iSiS)
            liSiS = np.log(np.linalg.det(iSiS))
            mij = iSmi + iSmj
            iiSSj = np.sum(mij * np.matmul(iiSiS, mij), axis=(1, 2))
            phi = miiSmi + mjiSmj - iiSSj + ldS[:, i] + ldS[:, j] + liSiS
            T2 += 2 * pps[i] * pps[j] * np.exp(-0.5 * phi)
    T2 = E / 2 * np.log(2 * np.pi) - np.log(T2)
    return T2 - T1


CRITERIA = {'HR': HR, 'BH': BH, 'BF': BF, 'AW': AW, 'JR': JR}
