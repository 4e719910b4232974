"""Oracle: CTC negative log-likelihood and its gradient w.r.t. the logits (numpy float64).

Restates what espresso/criterions/ctc_loss.py:59-103 computes: fp32 log_softmax
(speech_transformer_encoder_model.py:141-150) followed by F.ctc_loss(blank, reduction="sum",
zero_infinity) -- ATen's alpha/beta recursion (Graves et al. 2006, eq. 6-16).  Test infrastructure only.
"""
import numpy as np


def _lse(*xs):
    m = max(xs)
    if m == -np.inf:
        return -np.inf
    return m + np.log(sum(np.exp(x - m) for x in xs))


def ctc_loss_and_grad(logits, in_len, target, blank, zero_infinity=True):
    """logits [T, V] (any float dtype), target 1-D ints.  Returns (nll, dnll/dlogits [T, V]); rows t >= in_len get 0."""
    x = np.asarray(logits, dtype=np.float64)
    T_all, V = x.shape
    T = int(in_len)
    lp = x - (x.max(axis=1, keepdims=True) + np.log(np.exp(x - x.max(axis=1, keepdims=True)).sum(axis=1, keepdims=True)))
    U = len(target)
    S = 2 * U + 1
    ext = [blank if s % 2 == 0 else int(target[s // 2]) for s in range(S)]
    grad = np.zeros_like(x)
    if T == 0:
        nll = 0.0 if U == 0 else np.inf
        return (0.0 if (zero_infinity and np.isinf(nll)) else nll), grad
    alpha = np.full((T, S), -np.inf)
    beta = np.full((T, S), -np.inf)
    alpha[0, 0] = lp[0, ext[0]]
    if S > 1:
        alpha[0, 1] = lp[0, ext[1]]
    for t in range(1, T):
        for s in range(S):
            terms = [alpha[t - 1, s]]
            if s >= 1:
                terms.append(alpha[t - 1, s - 1])
            if s >= 2 and s % 2 == 1 and ext[s] != ext[s - 2]:
                terms.append(alpha[t - 1, s - 2])
            alpha[t, s] = _lse(*terms) + lp[t, ext[s]]
    ll = _lse(alpha[T - 1, S - 1], alpha[T - 1, S - 2]) if S > 1 else alpha[T - 1, 0]
    nll = -ll
    if np.isinf(nll):
        return (0.0 if zero_infinity else nll), grad
    beta[T - 1, S - 1] = lp[T - 1, ext[S - 1]]
    if S > 1:
        beta[T - 1, S - 2] = lp[T - 1, ext[S - 2]]
    for t in range(T - 2, -1, -1):
        for s in range(S):
            terms = [beta[t + 1, s]]
            if s + 1 < S:
                terms.append(beta[t + 1, s + 1])
            if s + 2 < S and s % 2 == 1 and ext[s] != ext[s + 2]:
                terms.append(beta[t + 1, s + 2])
            beta[t, s] = _lse(*terms) + lp[t, ext[s]]
    occ = np.zeros((T, V))
    for t in range(T):
        for s in range(S):
            ab = alpha[t, s] + beta[t, s]
            if ab > -np.inf:
                occ[t, ext[s]] += np.exp(ab - lp[t, ext[s]] + nll)
    grad[:T] = np.exp(lp[:T]) - occ
    return nll, grad
