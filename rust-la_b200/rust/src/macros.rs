// The `m!` macro -- same grammar and expansion as the reference's src/macros.rs:4-42 (host-side only: it builds a
// Vec and calls Matrix::new).  `m!(1.0, 2.0; 3.0, 4.0)` => Matrix::new(2, 2, vec![1.0, 2.0, 3.0, 4.0]).
#[macro_export]
macro_rules! m {
    ( $( $( $x:expr ),+ );+ ) => {{
        let mut data = Vec::new();
        let mut rows = 0usize;
        $(
            rows += 1;
            $( data.push($x); )+
        )+
        let cols = data.len() / rows;
        $crate::Matrix::new(rows, cols, data)
    }};
}
