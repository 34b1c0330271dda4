// oracle/bessel_shim.cpp -- TEST INFRASTRUCTURE.  The reference's PIC method calls libstdc++'s
// std::cyl_bessel_j / std::cyl_bessel_i (include/solver_pic.h:88,266,389 of ssskkkky/EMME); the
// plain-C restatement reaches the same two library functions through this shim.
#include <cmath>
extern "C" double emme_shim_cyl_bessel_j(double nu, double x) { return std::cyl_bessel_j(nu, x); }
extern "C" double emme_shim_cyl_bessel_i(double nu, double x) { return std::cyl_bessel_i(nu, x); }
