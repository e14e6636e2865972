/*
 * cpu_ref.h — ORACLE (test infrastructure, never shipped, never on the product path).
 *
 * A plain-C CPU restatement of the arrow-go compute hot path.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library.  Every function cites the reference lines it follows (paths relative
 * to the arrow-go tree, commit b3dacd2a).
 *
 * Pinning: tests/test_oracle_*.py check this restatement against
 *   (1) the reference's own instruction stream, oracle/_ref/libarrowgo_ref.so (assembled
 *       from the reference's checked-in _lib .s files; oracle/Makefile) for every function
 *       that has a native counterpart (sum, arithmetic, comparisons, aligned bitmap ops,
 *       numeric casts, integer min/max),
 *   (2) the literal known-answer vectors of the reference's Go tests, restated in
 *       tests/golden/ (arithmetic_test.go, scalar_compare_test.go, scalar_bool_test.go,
 *       vector_selection_test.go, arrow/math/{float64,int64,uint64}_test.go, bitmaps_test.go,
 *       cast_test.go, vector_cumulative_test.go),
 *   (3) pyarrow (an independent implementation of the same Arrow semantics) for the
 *       Go-only logic (filter, take, Kleene).
 */
#ifndef CPU_REF_H
#define CPU_REF_H
#include <stddef.h>
#include <stdint.h>

/* status codes mirror include/arrowgpu.h */
#define REF_OK 0
#define REF_ERR_INVALID 1
#define REF_ERR_INDEX 2
#define REF_ERR_NOT_IMPLEMENTED 3
#define REF_ERR_TYPE 4
#define REF_NO_ERROR_POS INT64_MAX

/* arrow/math */
double   ref_sum_f64_avx2_order(const double* buf, size_t n);   /* float64_avx2_amd64.s association */
double   ref_sum_f64_sequential(const double* buf, size_t n);   /* float64.go:41-47 (noasm) */
int64_t  ref_sum_i64(const int64_t* buf, size_t n);
uint64_t ref_sum_u64(const uint64_t* buf, size_t n);

/* arithmetic (all slots) */
int ref_arith_binary(int type, int op, int shape, const void* l, const void* r, void* out, int64_t n);
void ref_arith_binary_native_abi(int type, int8_t op, const void* l, const void* r, void* out, int len);
int ref_arith_unary_same(int type, int op, const void* in, void* out, int64_t n);
int ref_arith_unary_diff(int itype, int otype, int op, const void* in, void* out, int64_t n);
/* checked integer arithmetic with ScalarBinaryNotNull / ScalarBinary slot semantics */
int ref_arith_checked(int type, int op, int shape,
                      const void* l, const uint8_t* lvalid, int64_t loff,
                      const void* r, const uint8_t* rvalid, int64_t roff,
                      void* out, int64_t n, int64_t* first_bad);

int ref_arith_unary_checked(int type, int op, const void* in, void* out, int64_t n, int64_t* first_bad);

/* cumulative_sum[_checked]; state = {8 value bytes, int64 encountered_null} carried between chunks */
int ref_cumulative_sum(int type, const void* in, const uint8_t* valid, int64_t voff, int64_t n, int skip_nulls, int checked,
                       void* out, uint8_t* out_valid, int64_t ooff, void* state, int64_t* null_count, int64_t* first_bad);

/* integer min/max */
int ref_min_max(int type, const void* in, int64_t n, void* min_out, void* max_out);

/* numeric casts (loop + safe-cast checks) */
int ref_cast_numeric(int itype, int otype, const void* in, const uint8_t* valid, int64_t voff, void* out, int64_t n,
                     int allow_int_overflow, int allow_float_truncate, int64_t* first_bad);

/* comparisons -> bitmap */
int ref_compare(int type, int cmp, int shape, const void* l, const void* r, uint8_t* out, int64_t n, int offset);

/* bitmaps */
int     ref_bitmap_op(int bitop, const uint8_t* l, int64_t loff, const uint8_t* r, int64_t roff,
                      uint8_t* out, int64_t ooff, int64_t n);
void    ref_bitmap_copy(const uint8_t* src, int64_t soff, int64_t n, uint8_t* dst, int64_t doff);
void    ref_bitmap_invert(const uint8_t* src, int64_t soff, int64_t n, uint8_t* dst, int64_t doff);
void    ref_bitmap_set(uint8_t* bits, int64_t off, int64_t n, int value);
int64_t ref_bitmap_popcount(const uint8_t* bits, int64_t off, int64_t n);
int     ref_kleene(int kop, const uint8_t* lvalid, const uint8_t* ldata, int64_t loff,
                   const uint8_t* rvalid, const uint8_t* rdata, int64_t roff,
                   uint8_t* out_valid, uint8_t* out_data, int64_t ooff, int64_t n);

/* selection */
int64_t ref_filter_output_size(const uint8_t* mask, const uint8_t* mvalid, int64_t moff, int64_t n, int null_selection);
int ref_filter_primitive(int bit_width, const void* vals, const uint8_t* vvalid, int64_t voff,
                         const uint8_t* mask, const uint8_t* mvalid, int64_t moff, int64_t n,
                         int null_selection, void* out, uint8_t* out_valid, int64_t* out_len, int64_t* out_nulls);
int ref_take_indices(int index_width, const uint8_t* mask, const uint8_t* mvalid, int64_t moff, int64_t n,
                     int null_selection, void* out_idx, uint8_t* out_valid, int64_t* out_len);
int ref_take_primitive(int bit_width, const void* vals, const uint8_t* vvalid, int64_t voff, int64_t vlen,
                       int idx_width, int idx_signed, const void* idx, const uint8_t* ivalid, int64_t ioff,
                       int64_t n, int bounds_check, void* out, uint8_t* out_valid,
                       int64_t* out_nulls, int64_t* bad_pos, int64_t* bad_index);

/* parity helpers (same definitions as the device versions) */
int ref_is_in(int bit_width, const void* vals, const uint8_t* valid, int64_t off, int64_t n, const void* set, const uint8_t* set_valid,
              int64_t set_off, int64_t set_n, int null_behavior, uint8_t* out_data, uint8_t* out_valid, int64_t* out_nulls);
int ref_unique(int bit_width, const void* vals, const uint8_t* valid, int64_t off, int64_t n, void* out, uint8_t* out_valid, int64_t* out_len,
               int64_t* out_nulls);
int ref_sort_indices(int type, const void* vals, const uint8_t* valid, int64_t voff, int64_t n, int order, int null_placement, uint64_t* out,
                     int64_t* null_count, int64_t* nan_count);
uint64_t ref_checksum64(const void* buf, size_t n_words);
void     ref_generate(int kind, uint64_t seed, int64_t lo, int64_t hi, void* out, size_t n);

#endif
