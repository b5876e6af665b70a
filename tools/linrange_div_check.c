#include <stdio.h>
#include <math.h>
#include <stdlib.h>
// is  q = fma(fma(-q0, d, i), r, q0), q0 = i*r, r = RN(1/d)  equal to RN(i/d) for all integers 0 <= i <= d <= dmax ?
int main(int argc, char **argv) {
  long dmin = atol(argv[1]), dmax = atol(argv[2]);
  long bad = 0;
  #pragma omp parallel for schedule(dynamic, 64) reduction(+:bad)
  for (long d = dmin; d <= dmax; ++d) {
    const double fd = (double)d, r = 1.0 / fd;
    for (long i = 0; i <= d; ++i) {
      const double fi = (double)i;
      const double q0 = fi * r;
      const double rem = fma(-q0, fd, fi);
      const double q = fma(rem, r, q0);
      if (q != fi / fd) ++bad;
    }
  }
  printf("d in [%ld, %ld]: mismatches %ld\n", dmin, dmax, bad);
  return bad != 0;
}
