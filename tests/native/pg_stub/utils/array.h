/* TEST STAND-IN, see ../postgres.h: a one-dimensional array without the
 * varlena header games of the real ArrayType */
#ifndef PG_STUB_ARRAY_H
#define PG_STUB_ARRAY_H
typedef struct ArrayType
{
    int     ndim;
    int     dims[1];
    bool    hasnull;
    Oid     elemtype;
    char    data[];         /* 8-byte aligned by the members above */
} ArrayType;
#define ARR_NDIM(a)         ((a)->ndim)
#define ARR_DIMS(a)         ((a)->dims)
#define ARR_HASNULL(a)      ((a)->hasnull)
#define ARR_ELEMTYPE(a)     ((a)->elemtype)
#define ARR_DATA_PTR(a)     ((a)->data)
#define ARR_STUB_SIZE(a)    (sizeof(ArrayType) + 8 * (size_t) (a)->dims[0])
static inline ArrayType *pg_stub_array_copy(ArrayType *a)
{
    ArrayType *c = (ArrayType *) malloc(ARR_STUB_SIZE(a));  /* palloc: freed with the context */
    memcpy(c, a, ARR_STUB_SIZE(a));
    return c;
}
#define PG_GETARG_ARRAYTYPE_P(n)        ((ArrayType *) fcinfo->arg[n])
#define PG_GETARG_ARRAYTYPE_P_COPY(n)   pg_stub_array_copy((ArrayType *) fcinfo->arg[n])
#define PG_RETURN_ARRAYTYPE_P(a)        return PointerGetDatum(a)
#endif
