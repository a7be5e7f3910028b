// Faithful arithmetic mode: compiled with -fmad=false so that no multiply-add is contracted and every
// double operation rounds exactly as in the reference's operation order (gfortran -O3 on x86-64 emits
// no FMA).  See transport.cuh.
#define ARTES_FAITHFUL 1
#define ARTES_NS faithful
#include "transport.cuh"
#include "launchers.inc"
