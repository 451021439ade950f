"""Dense CPU port of the reference's step-4 graph -- TEST / BASELINE INFRASTRUCTURE ONLY.

The reference feeds a dense multi-hot X [N, n_genes] float32 to two tf.matmul's and lets TF1 autodiff
and ApplyAdam do the rest (/root/reference/G2Vec.py:231-251,262-267).  TensorFlow is not installed, so
the timed CPU baseline executes the same DENSE formulation with torch-CPU matmuls on all host threads:
one "epoch" = optimizer step on the training rows + accuracy on validation rows + accuracy on training
rows, exactly the three session runs of G2Vec.py:264-267.  Used by bench.py (cpu_baseline,
--impl reference) and cross-checked against the sparse oracle in tests/test_oracle_cbow.py.
"""
import numpy as np
import torch


def densify(rowptr, gene, label, idx, V):
    X = torch.zeros((len(idx), V), dtype=torch.float32)
    for r, n in enumerate(idx):
        X[r, torch.from_numpy(np.asarray(gene[rowptr[n]:rowptr[n + 1]], dtype=np.int64))] = 1.0
    y = torch.from_numpy(np.asarray(label)[idx].astype(np.float32)).reshape(-1, 1)
    return X, y


class DenseCbow:
    def __init__(self, W_ih0, W_ho0, lr, beta1=0.9, beta2=0.999, eps=1e-8):
        self.W = torch.from_numpy(np.array(W_ih0, dtype=np.float32, copy=True))
        self.Wo = torch.from_numpy(np.array(W_ho0, dtype=np.float32, copy=True)).reshape(-1, 1)
        self.st = [torch.zeros_like(self.W), torch.zeros_like(self.W), torch.zeros_like(self.Wo), torch.zeros_like(self.Wo)]
        self.lr, self.b1, self.b2, self.eps = lr, beta1, beta2, eps
        self.b1p, self.b2p = np.float32(1), np.float32(1)

    def _adam(self, var, m, v, g, alpha):
        m.add_((g - m) * (1 - self.b1))
        v.add_((g * g - v) * (1 - self.b2))
        var.sub_((m * alpha) / (v.sqrt() + self.eps))

    def accuracy(self, X, y):
        O = (X @ self.W) @ self.Wo                                   # G2Vec.py:239-240
        return float(((O > 0).float() == y).float().mean())          # :249-251

    def train_step(self, X, y):
        H = X @ self.W
        O = H @ self.Wo
        dO = (torch.sigmoid(O) - y) / X.shape[0]                     # d mean(BCE) / dO   (:243)
        gWo = H.t() @ dO
        gW = X.t() @ (dO @ self.Wo.t())                               # dense X^T . dH, as autodiff does
        self.b1p = np.float32(self.b1p * np.float32(self.b1)); self.b2p = np.float32(self.b2p * np.float32(self.b2))
        alpha = float(np.float32(self.lr) * np.sqrt(np.float32(1) - self.b2p) / (np.float32(1) - self.b1p))
        self._adam(self.W, self.st[0], self.st[1], gW, alpha)
        self._adam(self.Wo, self.st[2], self.st[3], gWo, alpha)

    def epoch(self, Xtr, ytr, Xva, yva):
        """One iteration of the loop at G2Vec.py:262-267."""
        self.train_step(Xtr, ytr)
        return self.accuracy(Xva, yva), self.accuracy(Xtr, ytr)
