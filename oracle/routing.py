"""Restatement of GPdoemd/utils.py:27-48 ``binary_dimensions`` (TEST INFRASTRUCTURE ONLY).

Rows are routed to the GP of their binary-variable combination.  Combination index = row index of
``vstack([ravel(g) for g in meshgrid(*[[-1,1]]*b)]).T`` whose signs match ``z_b - 0.5`` strictly.
The reference's ``np.vstack(map(...))`` is broken on numpy >= 2 (needs ``list()``); rows whose binary
value is exactly 0.5 match no combination and are silently dropped from J.
"""
import numpy as np


def sign_table(lb):
    B = np.meshgrid(*([[-1, 1]] * lb))
    return np.vstack([np.ravel(b) for b in B]).T


def binary_dimensions(Z, binary_variables):
    lb = len(binary_variables)
    if lb == 0:
        return [0], np.zeros(Z.shape[0])
    B = sign_table(lb)
    Zt = Z[:, binary_variables] - 0.5
    J = []
    for z in Zt:
        for j, b in enumerate(B):
            if np.all(b * z > 0):
                J.append(j)
                break
    return range(2 ** lb), np.array(J)
