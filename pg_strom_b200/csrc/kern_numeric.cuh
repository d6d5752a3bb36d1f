/*
 * kern_numeric.cuh - device NUMERIC (64-bit: 6-bit exp10 / sign / 57-bit
 * mantissa, opencl_numeric.h:141-162).  Filled in by a later step; programs
 * that need it are rejected by the planner until then.
 */
#ifndef KERN_NUMERIC_CUH
#define KERN_NUMERIC_CUH
#endif  /* KERN_NUMERIC_CUH */
