"""Posterior summaries (host numpy): shortest interval, MAP sample, histogram mode.

Same conventions as gptransits/stats.py:4-43: chain [walker, step, dim], posterior [step, walker].
"""
import numpy as np


def hpd(chain, level):
    flat = np.sort(chain.reshape(-1, chain.shape[2]), axis=0)
    width = int(np.floor(level * flat.shape[0]))
    lo = np.argmin(flat[width:] - flat[:-width], axis=0)
    cols = np.arange(flat.shape[1])
    return flat[lo, cols], flat[lo + width, cols]


def mapv(chain, posterior):
    step, walker = np.unravel_index(np.argmax(posterior), posterior.shape)
    return chain[walker, step, :]


def mode(chain, bins=50):
    flat = chain.reshape(-1, chain.shape[2])
    out = np.zeros(chain.shape[2])
    for i in range(out.size):
        counts, edges = np.histogram(flat[:, i], bins=bins)
        k = np.argmax(counts)
        out[i] = 0.5 * (edges[k] + edges[k + 1])
    return out
