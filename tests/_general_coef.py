"""PDE coefficients that are NOT constant / affine: a_ij(X) varies over the sample, c(X, u) depends on X and is quadratic
in u (so A(u) = c u is cubic).  The reference accepts any callables (src/training.py:30-41); these are the ones the golden
`extra/general_coef_cube_d3` was generated with (tests/golden/make_golden.py) and the tests re-create."""


def func_a(X, i, j):
    if i == j:
        return 1.0 + 0.5 * X[..., 1 + i] ** 2
    return 0.2 * X[..., 1 + i] * X[..., 1 + j]


def func_c(X, u):
    return -(1.0 + 0.3 * X[..., 1:2]) * u * u
