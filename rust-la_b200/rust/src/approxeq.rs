// Approximate comparison used by the reference's tests (`approx_eq` on scalars and, through matrix.rs, on matrices).
// Contract mirrored from the reference (src/approxeq.rs:12-47): an ABSOLUTE tolerance, strict `<`, default 1e-6 for both
// float widths; the trait keeps the reference's name and method set so user code and the crate's other modules compile
// unchanged.  The two implementations are generated from one macro.
pub trait ApproxEq<Eps> {
    fn approx_epsilon() -> Eps;
    fn approx_eq(&self, other: &Self) -> bool;
    fn approx_eq_eps(&self, other: &Self, approx_epsilon: &Eps) -> bool;
}

/// Default absolute tolerance of `approx_eq`, shared by f32 and f64.
pub const DEFAULT_ABS_TOLERANCE: f64 = 1.0e-6;

macro_rules! absolute_tolerance_impl {
    ($($float:ty),+) => {$(
        impl ApproxEq<$float> for $float {
            #[inline]
            fn approx_epsilon() -> $float {
                DEFAULT_ABS_TOLERANCE as $float
            }
            #[inline]
            fn approx_eq(&self, other: &$float) -> bool {
                let tolerance = <$float as ApproxEq<$float>>::approx_epsilon();
                self.approx_eq_eps(other, &tolerance)
            }
            #[inline]
            fn approx_eq_eps(&self, other: &$float, approx_epsilon: &$float) -> bool {
                let distance = if *self >= *other { *self - *other } else { *other - *self };
                distance < *approx_epsilon // NaN on either side compares false, as `(a - b).abs() < eps` does
            }
        }
    )+};
}
absolute_tolerance_impl!(f32, f64);
