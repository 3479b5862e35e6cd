"""Restatement of GPdoemd/marginal/taylor_first_order.py:27-43 and taylor_second_order.py:27-60
(TEST INFRASTRUCTURE ONLY).  Works with any duck-typed model exposing ``num_outputs, dim_p, Sigma,
predict, d_mu_d_p`` (+ ``d2_mu_d_p2, d2_s2_d_p2`` for second order)."""
import numpy as np


def taylor_first_order(model, xnew):
    N, E, D = len(xnew), model.num_outputs, model.dim_p
    M, s2 = model.predict(xnew)
    assert M.shape == (N, E) and s2.shape == (N, E)
    dmu = np.zeros((N, E, D))
    for e in range(E):
        dmu[:, e] = model.d_mu_d_p(e, xnew)
    S = np.zeros((N, E, E))
    for n in range(N):                                  # :37-39, same association order
        S[n] = np.diag(s2[n]) + np.matmul(dmu[n], np.matmul(model.Sigma, dmu[n].T))
    for e in range(E):                                  # :41-42
        S[:, e, e] = np.maximum(1e-15, S[:, e, e])
    return M, S


def taylor_second_order(model, xnew):
    N, E, D = len(xnew), model.num_outputs, model.dim_p
    M, s2 = model.predict(xnew)
    assert M.shape == (N, E) and s2.shape == (N, E)
    dmu = np.zeros((N, E, D))
    ddmuA = np.zeros((N, E, D, D))
    for e in range(E):
        dmu[:, e] = model.d_mu_d_p(e, xnew)
        ddmu = model.d2_mu_d_p2(e, xnew)
        dds2 = model.d2_s2_d_p2(e, xnew)
        ddmuA[:, e] = np.matmul(ddmu, model.Sigma)                              # :40
        M[:, e] += 0.5 * np.trace(ddmuA[:, e], axis1=1, axis2=2)                # :44
        s2[:, e] += 0.5 * np.sum(dds2 * model.Sigma, axis=(1, 2))               # :42,45
    S = np.zeros((N, E, E))
    for n in range(N):
        S[n] = np.diag(s2[n]) + np.matmul(dmu[n], np.matmul(model.Sigma, dmu[n].T))
        for e1 in range(E):
            for e2 in range(e1, E):
                S[n, e1, e2] += 0.5 * np.sum(ddmuA[n, e1] * ddmuA[n, e2].T)     # :53-55
                if e1 != e2:
                    S[n, e2, e1] = S[n, e1, e2]
            S[n, e1, e1] = np.maximum(1e-15, S[n, e1, e1])                      # :59
    return M, S
