//! print.rs — the reference pretty-prints the expression tree (src/print.rs); with opaque handles only values are
//! printable (see `impl Display for Function` in function.rs and `Matrix::print`). Kept so that `mod print;` in an
//! unchanged main.rs resolves.
