"""Small linear-algebra helpers used by the parameter containers (host side, numpy).
Behavioural reference: sgmcmc_ssm/_utils.py:17-183."""
import functools

import numpy as np
import scipy.stats


@functools.lru_cache(maxsize=None)
def _tril_indices(n):
    return np.tril_indices(n)


def tril_indices_from(mat):
    """np.tril_indices_from with the index arrays cached per size (the parameter containers call it on every
    gradient / projection; for the 1 x 1 matrices of this path the index construction dominated)."""
    return _tril_indices(np.shape(mat)[0])


def tril_vector_to_mat(vec):
    """Lower-triangular matrix from its row-major packed vector (_utils.py:134-139)."""
    vec = np.atleast_1d(vec)
    if len(vec) == 1:                      # the n = 1 models of the PF path: skip the index machinery
        return np.array(vec, dtype=float).reshape(1, 1)
    n = int(np.sqrt(len(vec) * 2))
    mat = np.zeros((n, n), dtype=float)
    mat[_tril_indices(n)] = vec
    return mat


def pos_def_mat_inv(mat):
    """Inverse of a symmetric positive-definite matrix (_utils.py:99-116)."""
    if np.isscalar(mat) or np.size(mat) == 1:
        return np.asarray(mat, dtype=float) ** -1
    L = np.linalg.cholesky(mat)
    Linv = np.linalg.inv(L)
    return Linv.T.dot(Linv)


def array_wishart_rvs(df, scale):
    """scipy wishart draw that is always a 2-d array (_utils.py:27-35)."""
    if np.size(scale) == 1:
        return np.array([[scipy.stats.wishart(df=df, scale=scale).rvs()]])
    return scipy.stats.wishart(df=df, scale=scale).rvs()


def matrix_normal_logpdf(X, mean, Lrowprec, Lcolprec):
    """Matrix-normal log density from Cholesky factors of the precisions (_utils.py:58-77)."""
    n, m = np.shape(X)
    out = -0.5 * n * m * np.log(2 * np.pi)
    out += -0.5 * np.sum(np.dot(Lrowprec.T, np.dot(X - mean, Lcolprec)) ** 2)
    out += m * np.sum(np.log(np.diag(Lrowprec)))
    out += n * np.sum(np.log(np.diag(Lcolprec)))
    return out


def varp_stability_projection(A, eigenvalue_cutoff=0.9999, var_name="A", logger=None):
    """Shrink a VAR(p) coefficient matrix so its companion spectral radius is <= cutoff
    (_utils.py:149-172).  Modifies and returns A."""
    m, mp = np.shape(A)
    p = mp // m
    if m > 1 or p > 1:
        F = np.concatenate([A, np.eye(N=m * (p - 1), M=m * p)])
        rho = np.max(np.abs(np.linalg.eig(F)[0]))
        if rho > eigenvalue_cutoff:
            for ii in range(p):
                A[:, m * ii:m * (ii + 1)] *= (eigenvalue_cutoff / rho) ** (ii + 1)
    else:
        rho = np.abs(A[0, 0])
        if rho > eigenvalue_cutoff:
            if logger is not None:
                logger.info("Thresholding |{2}|: {0} > {1}".format(rho, eigenvalue_cutoff, var_name))
            A *= (eigenvalue_cutoff / rho)
    return A


def var_stationary_precision(Qinv, A, num_iters=50):
    """Fixed-point iteration for the stationary precision of a VAR(1) (_utils.py:175-183)."""
    precision = Qinv
    QinvA = np.dot(Qinv, A)
    AtQinvA = np.dot(A.T, QinvA)
    for _ in range(num_iters):
        precision = Qinv - np.dot(QinvA, np.linalg.solve(precision + AtQinvA, QinvA.T))
    return precision
