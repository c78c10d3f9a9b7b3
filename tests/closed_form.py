"""float64 numpy closed forms of the four channelizer paths -- a second,
independent opinion used by the tests to cross-check the oracle.  Tiny and slow on purpose."""
import numpy as np


def pfbch2_analysis(h, M, m, x):
    """Y_t[k] = (1/M)(-1)^{kt} sum_l h[l] x[N_t - l] e^{+j2pi k l/M},  N_t = (t+1)M/2 - 1"""
    h = np.asarray(h, np.float64)[: 2 * M * m]
    x = np.asarray(x, np.complex128)
    T = len(x) // (M // 2)
    L = len(h)
    k = np.arange(M)
    Y = np.zeros((T, M), np.complex128)
    E = np.exp(2j * np.pi * np.outer(k, np.arange(L)) / M)
    for t in range(T):
        N = (t + 1) * M // 2 - 1
        idx = N - np.arange(L)
        xs = np.where(idx >= 0, x[np.clip(idx, 0, None)], 0)
        Y[t] = (E @ (h * xs)) / M * ((-1.0) ** (k * t))
    return Y


def pfbch_analysis(h, M, p, x):
    """Y_s[c] = sum_l h[l] x[sM + M-1-l] e^{+j2pi c (l+1)/M}   (unscaled)"""
    h = np.asarray(h, np.float64)[: M * p]
    x = np.asarray(x, np.complex128)
    S = len(x) // M
    L = len(h)
    c = np.arange(M)
    E = np.exp(2j * np.pi * np.outer(c, np.arange(L) + 1) / M)
    Y = np.zeros((S, M), np.complex128)
    for s in range(S):
        idx = s * M + M - 1 - np.arange(L)
        xs = np.where(idx >= 0, x[np.clip(idx, 0, None)], 0)
        Y[s] = E @ (h * xs)
    return Y


def pfbch_synthesis(h, M, p, X):
    """y[sM+i] = sum_n h[i+nM] v_{s-n}[i],  v_s = unscaled IDFT of frame s"""
    h = np.asarray(h, np.float64)[: M * p]
    X = np.asarray(X, np.complex128).reshape(-1, M)
    v = np.fft.ifft(X, axis=1) * M
    S = X.shape[0]
    y = np.zeros((S, M), np.complex128)
    for s in range(S):
        for n in range(p):
            if s - n >= 0:
                y[s] += h[np.arange(M) + n * M] * v[s - n]
    return y.reshape(-1)


def pfbch2_synthesis(h, M, m, X):
    """y[t M/2 + i] = sum_{q<4m} h[i + q M/2] v_{t-q}[i + (t odd ? M/2 : 0)],
    v_t = IDFT_unscaled(X_t) / M * (M/2)                         """
    h = np.asarray(h, np.float64)[: 2 * M * m]
    X = np.asarray(X, np.complex128).reshape(-1, M)
    M2 = M // 2
    v = np.fft.ifft(X, axis=1) * M2
    T = X.shape[0]
    y = np.zeros((T, M2), np.complex128)
    i = np.arange(M2)
    for t in range(T):
        j = i + (M2 if t % 2 == 1 else 0)
        for q in range(4 * m):
            if t - q >= 0:
                y[t] += h[i + q * M2] * v[t - q][j]
    return y.reshape(-1)
