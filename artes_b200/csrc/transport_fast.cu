// Fast arithmetic mode: FMA contraction allowed, CDF inversion through host-built prefix tables.
// Same physics and draw order as the faithful mode.  See transport.cuh.
#define ARTES_FAITHFUL 0
#define ARTES_NS fast
#include "transport.cuh"
#include "launchers.inc"
