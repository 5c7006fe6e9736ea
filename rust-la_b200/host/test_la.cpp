// test_la.cpp -- the reference's unit tests for the hot path, written against la.hpp (C++ mirror of the crate API).
// Sources: src/matrix/mod.rs:1479-1571, src/matrix/mmatrix.rs:234-259, src/decomp/lu.rs:281-375,
// src/decomp/cholesky.rs:146-183.
// Without a GPU every compute call throws la::LaError (no CPU fallback); `--require-gpu` turns that into a failure,
// otherwise the binary only checks the host-side contract (panics before FFI) and reports SKIP for the rest.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>

#include "la.hpp"

using la::LUDecomposition;
using la::Matrix;
using la::Panic;
typedef Matrix<double> Md;

static int failures = 0, skipped = 0, passed = 0;
static bool require_gpu = false;

static void run(const char* name, const std::function<void()>& body) {
  try {
    body();
    ++passed;
    printf("ok      %s\n", name);
  } catch (const la::LaError& e) {
    if (e.status == LA_ERR_NO_DEVICE && !require_gpu) {
      ++skipped;
      printf("SKIP    %s (%s)\n", name, e.what());
    } else {
      ++failures;
      printf("FAILED  %s: %s\n", name, e.what());
    }
  } catch (const std::exception& e) {
    ++failures;
    printf("FAILED  %s: %s\n", name, e.what());
  }
}
#define CHECK(c) do { if (!(c)) throw std::runtime_error("check failed: " #c); } while (0)
template <typename F> static void should_panic(F f) {
  try { f(); } catch (const Panic&) { return; }
  throw std::runtime_error("expected a panic");
}

int main(int argc, char** argv) {
  require_gpu = argc > 1 && !strcmp(argv[1], "--require-gpu");

  run("test_mul (mod.rs:1479)", [] {
    auto m1 = la::m<int64_t>({{1, 2}, {3, 4}});
    auto m2 = la::m<int64_t>({{3, 4}, {5, 6}});
    CHECK((m1 * m2).get_data() == std::vector<int64_t>({13, 16, 29, 36}));
    CHECK((la::m<double>({{1, 2}, {3, 4}}) * la::m<double>({{3, 4}, {5, 6}})).get_data() == std::vector<double>({13, 16, 29, 36}));
  });
  run("test_mul_incompatible (mod.rs:1486)", [] {
    should_panic([] { la::m<int64_t>({{1, 2}, {3, 4}}) * la::m<int64_t>({{1, 2}, {3, 4}, {5, 6}}); });
  });
  run("test_mmul (mmatrix.rs:234)", [] {
    auto a = la::m<int64_t>({{1, 2}, {3, 4}}), b = la::m<int64_t>({{3, 4}, {5, 6}}), c = la::m<int64_t>({{0, 0}, {0, 0}});
    a.mmul(b, c);
    CHECK(c.get_data() == std::vector<int64_t>({13, 16, 29, 36}));
  });
  run("test_mmul_incompatible (mmatrix.rs:243-259)", [] {
    auto a = la::m<int64_t>({{1, 2}, {3, 4}}), b = la::m<int64_t>({{3, 4}, {5, 6}});
    should_panic([&] { auto d = la::m<int64_t>({{0, 0, 0}, {0, 0, 0}}); a.mmul(b, d); });
    should_panic([&] { auto d = la::m<int64_t>({{0, 0}, {0, 0}, {0, 0}}); a.mmul(b, d); });
  });
  auto lu_pa = [](Md a) {
    LUDecomposition<double> lu(a);
    CHECK(lu.get_l() * lu.get_u() == lu.get_p() * a);  // exact ==
  };
  run("test_lu_square (lu.rs:281)", [&] { lu_pa(la::m<double>({{1, 2, 0}, {3, 6, -1}, {1, 2, 1}})); });
  run("test_lu2_m_over_n (lu.rs:291)", [&] { lu_pa(la::m<double>({{1, 2}, {3, 4}, {5, 6}})); });
  run("test_lu2_m_under_n (lu.rs:301)", [&] { lu_pa(la::m<double>({{1, 2, 3}, {4, 5, 6}})); });
  run("lu_solve_test (lu.rs:311)", [] {
    LUDecomposition<double> lu(la::m<double>({{2, 1, 0}, {1, 1, 0}, {0, 0, 1}}));
    CHECK(lu.solve(la::m<double>({{1}, {2}, {3}}))->approx_eq(la::m<double>({{-1}, {3}, {3}})));
  });
  run("lu_solve_test_incompatible (lu.rs:319)", [] {
    LUDecomposition<double> lu(la::m<double>({{2, 1, 0}, {1, 1, 0}, {0, 0, 1}}));
    should_panic([&] { lu.solve(la::m<double>({{1}, {2}, {3}, {4}})); });
  });
  run("lu_solve_test_singular (lu.rs:328)", [] {
    LUDecomposition<double> lu(la::m<double>({{2, 6}, {1, 3}}));
    CHECK(!lu.solve(la::m<double>({{1}, {2}})).has_value());
  });
  run("lu_is_singular_test (lu.rs:336)", [] {
    CHECK(LUDecomposition<double>(la::m<double>({{2, 6}, {1, 3}})).is_singular());
    CHECK(!LUDecomposition<double>(la::m<double>({{2, 6}, {1, 4}})).is_singular());
    CHECK(LUDecomposition<double>(la::m<double>({{4, 8}, {3, 4}})).is_non_singular());
    CHECK(!LUDecomposition<double>(la::m<double>({{4, 6}, {2, 3}})).is_non_singular());
  });
  run("lu_det_test (lu.rs:358)", [] {
    CHECK(LUDecomposition<double>(la::m<double>({{4, 8}, {3, 4}})).det() == -8.0);
    CHECK(LUDecomposition<double>(la::m<double>({{4, 8}, {2, 4}})).det() == 0.0);
  });
  run("lu_det_test_not_square (lu.rs:369)", [] {
    LUDecomposition<double> lu(la::m<double>({{1, 2, 3}, {4, 5, 6}}));
    should_panic([&] { lu.det(); });
  });
  run("test_det / test_solve / test_inverse (mod.rs:1506-1546)", [] {
    auto a = la::m<double>({{6, -7, 10}, {0, 3, -1}, {0, 5, -7}});
    CHECK(a.det() == -96.0);
    auto s = la::m<double>({{1, 1, 1}, {1, -1, 4}, {2, 3, -5}});
    CHECK(*s.solve(la::m<double>({{3}, {4}, {0}})) == la::m<double>({{1}, {1}, {1}}));
    auto inv = a.inverse();
    CHECK(inv.has_value());
    CHECK((a * *inv).approx_eq(Md::id(3, 3)));
    CHECK(!la::m<double>({{2, 6}, {1, 3}}).inverse().has_value());
  });
  run("test_is_singular (mod.rs:1554)", [] {
    CHECK(la::m<double>({{2, 6}, {1, 3}}).is_singular());
    CHECK(la::m<double>({{2, 6}, {6, 3}}).is_non_singular());
  });
  run("f32 instance", [] {
    auto a = la::m<float>({{4, 8}, {3, 4}});
    CHECK(a.det() == -8.0f);
    CHECK((a * Matrix<float>::id(2, 2)) == a);
  });
  run("cholesky (cholesky.rs:146-183)", [] {
    typedef la::CholeskyDecomposition<double> Chol;
    auto a = la::m<double>({{4, 12, -16}, {12, 37, -43}, {-16, -43, 98}});
    auto c = Chol::make(a);
    CHECK(c.has_value());
    auto l = c->get_l();
    CHECK(l.get_data() == std::vector<double>({2, 0, 0, 6, 1, 0, -8, 5, 3}));
    CHECK(l * l.t() == a);
    CHECK(!Chol::make(la::m<double>({{4, 12, -16}, {12, 37, 43}, {-16, 43, 98}})).has_value());  // not positive definite
    CHECK(!Chol::make(la::m<double>({{4, 12, -16}, {12, 37, 43}})).has_value());                  // not square
    auto s = Chol::make(la::m<double>({{2, 1, 0}, {1, 1, 0}, {0, 0, 1}}));
    CHECK(s.has_value());
    CHECK(s->solve(la::m<double>({{1}, {2}, {3}})).approx_eq(la::m<double>({{-1}, {3}, {3}})));
    should_panic([&] { s->solve(la::m<double>({{1}, {2}, {3}, {4}})); });
  });
  printf("%d passed, %d skipped, %d failed\n", passed, skipped, failures);
  return failures ? 1 : 0;
}
