"""The reference's own test-suite (src/main.rs:46-166), transliterated over the `Api` surface so the
same cases run against the CPU oracle and against the CUDA product. Each case yields the functions the
`test!` macro builds (scalar param x=0.5 and 2x2 matrix param m=0.5; for binary ops all 8
{const,param} x {scalar,matrix} x {lhs,rhs} combinations, main.rs:70-77), the learning rate and the
`all_less_than` goal. `minimize(&HashMap::new(), lr, 10, 1)` then `assert!(f.all_less_than(goal))`.
"""
START = 0.5  # main.rs:44


def _unary_case(api, build):
    F = api.Function
    x = F.scalar_param("x", START)
    m = F.single_val_param("m", [2, 2], START)
    return [build(x), build(m)]


def _binary_case(api, build, const_val):
    F = api.Function
    cx = F.scalar(const_val)
    cm = F.single_val_matrix(2, 2, const_val)
    px = F.scalar_param("x", START)
    pm = F.single_val_param("m", [2, 2], START)
    c = lambda f: f.clone()
    return [build(c(cx), c(px)), build(c(cx), c(pm)), build(c(cm), c(px)), build(c(cm), c(pm)),
            build(c(px), c(cx)), build(c(px), c(cm)), build(c(pm), c(cx)), build(c(pm), c(cm))]


def _broadcast(api):
    F = api.Function

    def build(x):
        a = F.scalar(10.)
        b = F.matrix(2, 2, [1.2, -1., 0., 1.5])
        return (x * a) + b
    return build


def _complex(api):
    F = api.Function

    def build(x):
        a = F.scalar(2.)
        b = F.matrix(2, 2, [1.2, -1., 0., 1.5])
        c = F.scalar(10.)
        return (((x - a) * (x + b) - c)).sq()
    return build


def _dot_case(api):
    F = api.Function
    x = F.param("m", api.Constant.single_val([2, 2], START))
    c = F.single_val_matrix(2, 2, 3.)
    return [api.dot(c, x), api.dot(x, c)]


# name -> (builder(api) -> [Function], lr, goal, main.rs lines)
CASES = {
    "single_param_test": (lambda api: _unary_case(api, lambda x: x), 1., -9.49, "85-87"),
    "neg_test": (lambda api: _unary_case(api, lambda x: -x), 1., -9.49, "89-91"),
    "sq_test": (lambda api: _unary_case(api, lambda x: x.sq()), 0.1, 0.01, "93-95"),
    "abs_test": (lambda api: _unary_case(api, lambda x: x.abs()), 0.1, 0.00001, "97-99"),
    "sigmoid_test": (lambda api: _unary_case(api, lambda x: x.sigmoid()), 1., -9.49, "101-103"),
    "tanh_test": (lambda api: _unary_case(api, lambda x: x.tanh()), 1., -0.97, "105-107"),
    "add_test": (lambda api: _binary_case(api, lambda x, y: x + y, 1.), 1., -7.4, "109-112"),
    "mul_test": (lambda api: _binary_case(api, lambda x, y: x * y, 2.), 0.1, -2.60, "114-117"),
    "sub_scalar_l_test": (lambda api: _unary_case(api, lambda x: api.Function.scalar(1.) - x), 1., -8.4, "119-121"),
    "sub_matrix_r_test": (lambda api: _unary_case(api, lambda x: x - api.Function.single_val_matrix(2, 2, 1.)), 1.,
                          -9.4, "123-126"),
    "sub_matrix_l_test": (lambda api: _unary_case(api, lambda x: api.Function.single_val_matrix(2, 2, 1.) - x), 1.,
                          -8.4, "128-131"),
    "broadcast_test": (lambda api: _unary_case(api, _broadcast(api)), 0.04, -29.4, "133-142"),
    "complex_test": (lambda api: _unary_case(api, _complex(api)), 0.04, 10.8, "144-154"),
    "dot_test": (_dot_case, 0.1, -29., "156-166"),
}


def run_case(api, name, trace=False):
    """Runs one reference test; returns [(function, per-iteration root values or None)]."""
    build, lr, goal, _ = CASES[name]
    out = []
    for f in build(api):
        tr = f.minimize({}, lr, 10, 1, trace=trace)
        out.append((f, tr))
    return out, goal
