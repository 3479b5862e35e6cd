"""Restatement of GPdoemd/transform.py:27-73 (TEST INFRASTRUCTURE ONLY)."""
import numpy as np


class BoxTransform:                      # transform.py:27-47
    def __init__(self, xmin, xmax):
        self.xmin, self.xmax = xmin, xmax
        self.diff = xmax - xmin

    def __call__(self, X, back=False):
        return self.xmin + X * self.diff if back else (X - self.xmin) / self.diff

    def var(self, X, back=False):
        return X * self.diff ** 2 if back else X / self.diff ** 2

    def cov(self, X, back=False):
        d = self.diff[:, None] * self.diff[None, :]
        return X * d if back else X / d


class MeanTransform:                     # transform.py:50-73
    def __init__(self, mean, std):
        self.mean, self.std = mean, std

    def __call__(self, X, back=False):
        return self.mean + X * self.std if back else (X - self.mean) / self.std

    def var(self, X, back=False):
        return X * self.std ** 2 if back else X / self.std ** 2

    def cov(self, X, back=False):
        d = self.std[:, None] * self.std[None, :]
        return X * d if back else X / d

    def prediction(self, M, S, back=False):
        Mt = self(M, back=back)
        return (Mt, self.var(S, back=back)) if S.ndim == 2 else (Mt, self.cov(S, back=back))
