"""Convergence diagnostics the north star adds to the reference (which has none):
Gelman-Rubin potential scale reduction R-hat (Gelman & Rubin 1992; Vrugt et al. 2009
use it as DREAM's stopping rule), computed on the device history."""
import numpy as np


def gelman_rubin(hist):
    """R-hat per dimension from a [T, n_chains, d] history (torch tensor or numpy):
    W = mean within-chain variance, B/T = variance of chain means,
    R = sqrt(((T-1)/T W + B/T) / W)."""
    try:
        import torch
        if isinstance(hist, torch.Tensor):
            T = hist.shape[0]
            cm = hist.mean(dim=0)
            W = hist.var(dim=0, unbiased=True).mean(dim=0)
            B_over_T = cm.var(dim=0, unbiased=True)
            return torch.sqrt(((T - 1.0) / T * W + B_over_T) / W).cpu().numpy()
    except ImportError:
        pass
    hist = np.asarray(hist)
    T = hist.shape[0]
    W = hist.var(axis=0, ddof=1).mean(axis=0)
    B_over_T = hist.mean(axis=0).var(axis=0, ddof=1)
    return np.sqrt(((T - 1.0) / T * W + B_over_T) / W)
