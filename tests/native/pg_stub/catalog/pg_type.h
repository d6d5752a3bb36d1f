/* TEST STAND-IN, see ../postgres.h */
#define INT8OID     20
#define FLOAT8OID   701
#define TEXTOID     25
#define BPCHAROID   1042
