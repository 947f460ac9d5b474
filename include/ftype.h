/* Element type switch of the approximateNN C API.
 *
 * Binary-compatible with the reference's ftype.h (/root/reference/ftype.h:3-9):
 * one library is built per element type; -DUSE_FLOAT selects float, otherwise
 * double.  i_ftype is the same-width signed integer used for bit comparisons. */
#ifndef FTYPE_H
#define FTYPE_H
#ifdef USE_FLOAT
#define ftype float
#define i_ftype int
#else
#define ftype double
#define i_ftype long
#endif
#endif
