/* TEST STAND-IN, see ../postgres.h (numeric_mul, int8_avg_accum, numeric_avg_accum
 * live in PostgreSQL; the glue's numeric wrappers are compiled out under the stub) */
