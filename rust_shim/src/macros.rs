// macros.rs — `dot!` exactly as user code spells it (reference src/macros.rs:51-64). The reference's `exec!`,
// `value!` and `value_mut!` macros drove its per-node interpreter and have no counterpart here: the walk is
// compiled once into a tape inside libcg_b200.so.
macro_rules! dot {
    ($f1:expr, $f2:expr) => {
        dot($f1, false, $f2, false)
    };
    (($f1:expr)^T, $f2:expr) => {
        dot($f1, true, $f2, false)
    };
    ($f1:expr, ($f2:expr)^T) => {
        dot($f1, false, $f2, true)
    };
    (($f1:expr)^T, ($f2:expr)^T) => {
        dot($f1, true, $f2, true)
    };
}
