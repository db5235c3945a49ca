"""Chain diagnostics consumed by Model.statistics (host numpy; not on the sampling path).

Same definitions and return conventions as gptransits/convergence.py:4-99, including the two the
reference is known to deviate from the textbook on: gelman_rubin returns
the tuple (V/W, V, W) without a square root, and gelman_brooks multiplies W^-1 and B/n elementwise.
chain is [nwalkers, niterations, nparameters].
"""
import numpy as np


def gelman_rubin(chain):
    m, n, _ = chain.shape
    walker_mean = chain.mean(axis=1)
    grand_mean = chain.mean(axis=(0, 1))
    between = ((walker_mean - grand_mean) ** 2).sum(axis=0) / (m - 1)              # B / n
    within = ((chain - walker_mean[:, None, :]) ** 2).sum(axis=(0, 1)) / (m * (n - 1))
    pooled = within * (n - 1) / n + between + between / m
    return pooled / within, pooled, within


def gelman_brooks(chain):
    m, n, _ = chain.shape
    dev_b = chain.mean(axis=1) - chain.mean(axis=(0, 1))
    between = dev_b.T @ dev_b / (m - 1)
    dev_w = chain - chain.mean(axis=1, keepdims=True)
    within = np.einsum("wnp,wnq->pq", dev_w, dev_w) / (m * (n - 1))
    eig = np.linalg.eigvals(np.linalg.inv(within) * between)   # elementwise, as the reference
    return (n - 1) / n + (1 + 1 / m) * eig


def geweke(chain, frac1=0.1, frac2=0.5, bins=10):
    nwalk, nit, npar = chain.shape
    starts = np.linspace(0, nit / 2, num=bins).astype(int)
    z = np.full([bins, nwalk, npar], 100.0)
    for b, s in enumerate(starts):
        sub = chain[:, s:, :]
        k1, k2 = int(frac1 * sub.shape[1]), int(frac2 * sub.shape[1])
        head, tail = sub[:, :k1, :], sub[:, k2:, :]
        with np.errstate(all="ignore"):
            z[b] = np.abs((head.mean(axis=1) - tail.mean(axis=1)) / np.sqrt(head.var(axis=1) + tail.var(axis=1)))
    return starts, z
