/*
 * cpu_ref.c — ORACLE: CPU restatement of arrow-go's compute hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under arrow_go_b200/ links, loads or calls this
 * file.  See cpu_ref.h for how it is pinned against the reference.
 *
 * Reference paths are relative to the arrow-go tree (commit b3dacd2a):
 *   K = arrow/compute/internal/kernels
 */
#include "cpu_ref.h"

#include <math.h>
#include <string.h>
#include <stdlib.h>

/* arrow.Type ids, arrow/datatype.go:36-72 */
enum { T_BOOL = 1, T_U8 = 2, T_I8 = 3, T_U16 = 4, T_I16 = 5, T_U32 = 6, T_I32 = 7, T_U64 = 8, T_I64 = 9, T_F32 = 11, T_F64 = 12 };
/* ArithmeticOp, K/base_arithmetic.go:37-82 */
enum { OP_ADD = 0, OP_SUB = 1, OP_MUL = 2, OP_DIV = 3, OP_ABS = 4, OP_NEG = 5, OP_SIGN = 20,
       OP_ADD_C = 21, OP_SUB_C = 22, OP_MUL_C = 23, OP_DIV_C = 24, OP_ABS_C = 25, OP_NEG_C = 26,
       OP_BIT_AND = 64, OP_BIT_OR = 65, OP_BIT_XOR = 66, OP_BIT_NOT = 67, OP_SHL = 68, OP_SHR = 69, OP_SHL_C = 70, OP_SHR_C = 71 };
enum { SH_AA = 0, SH_AS = 1, SH_SA = 2 };
enum { CMP_EQ = 0, CMP_NE = 1, CMP_GT = 2, CMP_GE = 3, CMP_LT = 4, CMP_LE = 5 };

/* ---- bit helpers: arrow/bitutil/bitutil.go:50-80 ------------------------------------ */
static inline int bit_is_set(const uint8_t* b, int64_t i) { return (b[i >> 3] >> (i & 7)) & 1; }
static inline void set_bit_to(uint8_t* b, int64_t i, int v) {
  /* K/_lib/scalar_comparison.cc:59-61 — only bit i of byte i/8 changes */
  b[i >> 3] = (uint8_t)((b[i >> 3] & ~(1u << (i & 7))) | ((unsigned)(v != 0) << (i & 7)));
}

/* ====================================================================================== *
 * arrow/math Sum
 * ====================================================================================== */

/* float64_avx2_amd64.s:36-43 (8 ymm accumulators zeroed), :86-164 (unrolled body: acc[y] +=
 * 4 lanes at x[32k + 4y ..]), combine :165-174:
 *   y1+=y5; y3+=y7; y0+=y4; y2+=y6; y0+=y2; y1+=y3; y0+=y1; xmm = lo128+hi128; hadd;
 * then the n&31 tail is added sequentially (:52-60).  n < 32 is purely sequential.
 * volatile stops the compiler from re-associating or contracting. */
double ref_sum_f64_avx2_order(const double* buf, size_t n) {
  volatile double acc[8][4];
  size_t body = n & ~(size_t)31;
  double s = 0.0;
  if (n > 31 && body != 0) {
    for (int y = 0; y < 8; ++y) for (int q = 0; q < 4; ++q) acc[y][q] = 0.0;
    for (size_t k = 0; k < body; k += 32)
      for (int y = 0; y < 8; ++y)
        for (int q = 0; q < 4; ++q) acc[y][q] = acc[y][q] + buf[k + 4 * y + q];
    volatile double v[4];
    for (int q = 0; q < 4; ++q) {
      volatile double t1 = acc[1][q] + acc[5][q];
      volatile double t3 = acc[3][q] + acc[7][q];
      volatile double t0 = acc[0][q] + acc[4][q];
      volatile double t2 = acc[2][q] + acc[6][q];
      volatile double u0 = t0 + t2;
      volatile double u1 = t1 + t3;
      v[q] = u0 + u1;
    }
    volatile double w0 = v[0] + v[2];
    volatile double w1 = v[1] + v[3];
    s = w0 + w1;
  } else {
    body = 0;
  }
  volatile double r = s;
  for (size_t i = body; i < n; ++i) r = r + buf[i];
  return r;
}

/* arrow/math/float64.go:41-47 */
double ref_sum_f64_sequential(const double* buf, size_t n) {
  volatile double acc = 0.0;
  for (size_t i = 0; i < n; ++i) acc = acc + buf[i];
  return acc;
}
/* arrow/math/int64.go:41-47, _lib/int64.c:21-27: wrapping */
int64_t ref_sum_i64(const int64_t* buf, size_t n) {
  uint64_t acc = 0;
  for (size_t i = 0; i < n; ++i) acc += (uint64_t)buf[i];
  return (int64_t)acc;
}
uint64_t ref_sum_u64(const uint64_t* buf, size_t n) {
  uint64_t acc = 0;
  for (size_t i = 0; i < n; ++i) acc += buf[i];
  return acc;
}

/* ====================================================================================== *
 * Arithmetic, all slots: K/_lib/base_arithmetic.cc:76-285
 * ====================================================================================== */
#define BIN_LOOP(T, EXPR)                                                               \
  do {                                                                                  \
    const T* L = (const T*)l; const T* R = (const T*)r; T* O = (T*)out;                 \
    for (int64_t i = 0; i < n; ++i) {                                                   \
      const T a = (shape == SH_SA) ? L[0] : L[i];                                       \
      const T b = (shape == SH_AS) ? R[0] : R[i];                                       \
      O[i] = (T)(EXPR);                                                                 \
    }                                                                                   \
  } while (0)

/* integers: arithmetic on the unsigned type of the same width (wraps; the low bits of
 * signed and unsigned add/sub/mul agree — base_arithmetic.cc:121-148 does the same for
 * Multiply via to_unsigned / uint32 promotion). */
#define BIN_INT(UT)                                                                     \
  switch (op) {                                                                         \
    case OP_ADD: case OP_ADD_C: BIN_LOOP(UT, (UT)(a + b)); return REF_OK;               \
    case OP_SUB: case OP_SUB_C: BIN_LOOP(UT, (UT)(a - b)); return REF_OK;               \
    case OP_MUL: case OP_MUL_C: BIN_LOOP(UT, (UT)((uint64_t)a * (uint64_t)b)); return REF_OK; \
    /* bitwiseKernelOp K/scalar_arithmetic.go:191-243: BitmapAnd/Or/Xor over the value buffers */ \
    case OP_BIT_AND: BIN_LOOP(UT, (UT)(a & b)); return REF_OK;                          \
    case OP_BIT_OR: BIN_LOOP(UT, (UT)(a | b)); return REF_OK;                           \
    case OP_BIT_XOR: BIN_LOOP(UT, (UT)(a ^ b)); return REF_OK;                          \
    default: return REF_ERR_NOT_IMPLEMENTED;                                            \
  }
#define BIN_FLT(FT)                                                                     \
  switch (op) {                                                                         \
    case OP_ADD: case OP_ADD_C: BIN_LOOP(FT, a + b); return REF_OK;                     \
    case OP_SUB: case OP_SUB_C: BIN_LOOP(FT, a - b); return REF_OK;                     \
    case OP_MUL: case OP_MUL_C: BIN_LOOP(FT, a * b); return REF_OK;                     \
    default: return REF_ERR_NOT_IMPLEMENTED;                                            \
  }

int ref_arith_binary(int type, int op, int shape, const void* l, const void* r, void* out, int64_t n) {
  switch (type) {
    case T_U8: case T_I8: BIN_INT(uint8_t)
    case T_U16: case T_I16: BIN_INT(uint16_t)
    case T_U32: case T_I32: BIN_INT(uint32_t)
    case T_U64: case T_I64: BIN_INT(uint64_t)
    case T_F32: BIN_FLT(float)
    case T_F64: BIN_FLT(double)
    default: return REF_ERR_TYPE;
  }
}

/* same loop behind the reference's native signature (base_arithmetic.cc:465: type, op, l, r, out, len) so the C
 * bench harness can time the restatement when oracle/_ref was never built */
void ref_arith_binary_native_abi(int type, int8_t op, const void* l, const void* r, void* out, int len) {
  (void)ref_arith_binary(type, op, 0, l, r, out, len);
}

/* unary: AbsoluteValue :160-176, Negate :194-205, NegateChecked :207-219 (unsigned -> 0),
 * Sign :221-234.  Float abs clears the sign bit (NaN payload preserved); float negate
 * is a sign flip (-x). */
#define UN_LOOP(TI, TO, EXPR)                                                           \
  do {                                                                                  \
    const TI* I = (const TI*)in; TO* O = (TO*)out;                                      \
    for (int64_t i = 0; i < n; ++i) { const TI x = I[i]; (void)x; O[i] = (TO)(EXPR); }           \
  } while (0)

#define UN_SINT(ST, UT)                                                                 \
  switch (op) {                                                                         \
    case OP_ABS: case OP_ABS_C: UN_LOOP(ST, ST, (ST)(((UT)x + (UT)(x >> (sizeof(ST) * 8 - 1))) ^ (UT)(x >> (sizeof(ST) * 8 - 1)))); return REF_OK; \
    case OP_NEG: case OP_NEG_C: UN_LOOP(ST, ST, (ST)(0 - (UT)x)); return REF_OK;        \
    case OP_SIGN: UN_LOOP(ST, ST, x > 0 ? 1 : (x ? -1 : 0)); return REF_OK;             \
    case OP_BIT_NOT: UN_LOOP(ST, ST, (ST)~x); return REF_OK;  /* bitwiseNot K/scalar_arithmetic.go:257-259 */ \
    default: return REF_ERR_NOT_IMPLEMENTED;                                            \
  }
#define UN_UINT(UT)                                                                     \
  switch (op) {                                                                         \
    case OP_ABS: case OP_ABS_C: UN_LOOP(UT, UT, x); return REF_OK;                      \
    case OP_NEG: UN_LOOP(UT, UT, (UT)(~x + 1)); return REF_OK;                          \
    case OP_NEG_C: UN_LOOP(UT, UT, 0); return REF_OK;                                   \
    case OP_SIGN: UN_LOOP(UT, UT, x > 0 ? 1 : 0); return REF_OK;                        \
    case OP_BIT_NOT: UN_LOOP(UT, UT, (UT)~x); return REF_OK;                            \
    default: return REF_ERR_NOT_IMPLEMENTED;                                            \
  }

static inline float f32_abs(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0x7fffffffu; memcpy(&x, &u, 4); return x; }
static inline double f64_abs(double x) { uint64_t u; memcpy(&u, &x, 8); u &= 0x7fffffffffffffffull; memcpy(&x, &u, 8); return x; }
static inline float f32_neg(float x) { uint32_t u; memcpy(&u, &x, 4); u ^= 0x80000000u; memcpy(&x, &u, 4); return x; }
static inline double f64_neg(double x) { uint64_t u; memcpy(&u, &x, 8); u ^= 0x8000000000000000ull; memcpy(&x, &u, 8); return x; }
/* Sign on floats: the source reads isnan(x) ? x : ... (base_arithmetic.cc:224-225) but the
 * reference's shipped AVX2/SSE4 objects were compiled with -funsafe-math-optimizations /
 * -fno-trapping-math (kernels/Makefile:23-27) and the NaN test is gone from the instruction
 * stream: sign(NaN) = +-1 by sign bit.  We follow the instruction stream (it is what runs on
 * amd64); the pure-Go fallback (base_arithmetic.go:427-441) returns NaN instead. */
#define SIGN_F(x) (((x) == 0) ? 0 : (signbit(x) ? -1 : 1))

int ref_arith_unary_same(int type, int op, const void* in, void* out, int64_t n) {
  switch (type) {
    case T_I8: UN_SINT(int8_t, uint8_t)
    case T_I16: UN_SINT(int16_t, uint16_t)
    case T_I32: UN_SINT(int32_t, uint32_t)
    case T_I64: UN_SINT(int64_t, uint64_t)
    case T_U8: UN_UINT(uint8_t)
    case T_U16: UN_UINT(uint16_t)
    case T_U32: UN_UINT(uint32_t)
    case T_U64: UN_UINT(uint64_t)
    case T_F32:
      switch (op) {
        case OP_ABS: case OP_ABS_C: UN_LOOP(float, float, f32_abs(x)); return REF_OK;
        case OP_NEG: case OP_NEG_C: UN_LOOP(float, float, f32_neg(x)); return REF_OK;
        case OP_SIGN: UN_LOOP(float, float, SIGN_F(x)); return REF_OK;
        default: return REF_ERR_NOT_IMPLEMENTED;
      }
    case T_F64:
      switch (op) {
        case OP_ABS: case OP_ABS_C: UN_LOOP(double, double, f64_abs(x)); return REF_OK;
        case OP_NEG: case OP_NEG_C: UN_LOOP(double, double, f64_neg(x)); return REF_OK;
        case OP_SIGN: UN_LOOP(double, double, SIGN_F(x)); return REF_OK;
        default: return REF_ERR_NOT_IMPLEMENTED;
      }
    default: return REF_ERR_TYPE;
  }
}

/* arithmetic_unary_diff_type: only SIGN (base_arithmetic.cc:427-438); result -1/0/1 (NaN
 * for float NaN when the output is float) converted to the output type. */
#define SIGN_OUT(TI, SEXPR)                                                             \
  switch (otype) {                                                                      \
    case T_U8: UN_LOOP(TI, uint8_t, SEXPR); return REF_OK;                              \
    case T_I8: UN_LOOP(TI, int8_t, SEXPR); return REF_OK;                               \
    case T_U16: UN_LOOP(TI, uint16_t, SEXPR); return REF_OK;                            \
    case T_I16: UN_LOOP(TI, int16_t, SEXPR); return REF_OK;                             \
    case T_U32: UN_LOOP(TI, uint32_t, SEXPR); return REF_OK;                            \
    case T_I32: UN_LOOP(TI, int32_t, SEXPR); return REF_OK;                             \
    case T_U64: UN_LOOP(TI, uint64_t, SEXPR); return REF_OK;                            \
    case T_I64: UN_LOOP(TI, int64_t, SEXPR); return REF_OK;                             \
    case T_F32: UN_LOOP(TI, float, SEXPR); return REF_OK;                               \
    case T_F64: UN_LOOP(TI, double, SEXPR); return REF_OK;                              \
    default: return REF_ERR_TYPE;                                                       \
  }
#define SIGN_S(x) ((x) > 0 ? 1 : ((x) ? -1 : 0))
#define SIGN_U(x) ((x) > 0 ? 1 : 0)

int ref_arith_unary_diff(int itype, int otype, int op, const void* in, void* out, int64_t n) {
  if (op != OP_SIGN) return REF_ERR_NOT_IMPLEMENTED;
  switch (itype) {
    case T_I8: SIGN_OUT(int8_t, SIGN_S(x))
    case T_I16: SIGN_OUT(int16_t, SIGN_S(x))
    case T_I32: SIGN_OUT(int32_t, SIGN_S(x))
    case T_I64: SIGN_OUT(int64_t, SIGN_S(x))
    case T_U8: SIGN_OUT(uint8_t, SIGN_U(x))
    case T_U16: SIGN_OUT(uint16_t, SIGN_U(x))
    case T_U32: SIGN_OUT(uint32_t, SIGN_U(x))
    case T_U64: SIGN_OUT(uint64_t, SIGN_U(x))
    default: return REF_ERR_NOT_IMPLEMENTED; /* float inputs: the Go side only uses same-type Sign */
  }
}

/* ====================================================================================== *
 * Checked integer arithmetic
 *   ADD_CHECKED  K/base_arithmetic.go:249-263   carry = ((a&b) | ((a|b) &^ out)) >> shiftBy ; carry > 0
 *   SUB_CHECKED  :264-278                        carry = ((^a&b) | (^(a^b) & out)) >> shiftBy ; carry > 0
 *     shiftBy = bits-1 for unsigned, bits-2 for signed, evaluated IN THE TYPE OutT — i.e. an
 *     arithmetic shift followed by a signed "> 0" for signed types.  (For signed types this
 *     flags exactly "carry into the sign bit without carry out of it"; we restate the formula,
 *     not the intent.)
 *   slots: ScalarBinaryNotNull, K/helpers.go:284-380 — visit valid∧valid slots, write the
 *     zero value into null slots; an all-null side leaves the output untouched.
 *   MUL_CHECKED  :279-286 + mulWithOverflow :84-108 under ScalarBinary (EVERY slot, null or not);
 *     an overflowing slot yields 0.
 *   DIV / DIV_CHECKED :154-161,287-294: b == 0 -> errDivByZero, result 0; else a / b (Go
 *     semantics: MinInt / -1 wraps to MinInt), ScalarBinaryNotNull.
 * ====================================================================================== */
#define CHK_SLOT_VALID(i)                                                               \
  ((!lvalid || shape == SH_SA || bit_is_set(lvalid, loff + (i))) &&                    \
   (!rvalid || shape == SH_AS || bit_is_set(rvalid, roff + (i))))

#define CHK_ADDSUB(ST, IS_SIGNED, CARRY_EXPR, RES_EXPR)                                 \
  do {                                                                                  \
    const ST* L = (const ST*)l; const ST* R = (const ST*)r; ST* O = (ST*)out;           \
    const int shift_by = (int)sizeof(ST) * 8 - 1 - (IS_SIGNED);                         \
    for (int64_t i = 0; i < n; ++i) {                                                   \
      if (!CHK_SLOT_VALID(i)) { O[i] = 0; continue; }                                   \
      const ST a = (shape == SH_SA) ? L[0] : L[i];                                      \
      const ST b = (shape == SH_AS) ? R[0] : R[i];                                      \
      const ST o = (ST)(RES_EXPR);                                                      \
      const ST carry = (ST)((ST)(CARRY_EXPR) >> shift_by);                              \
      if (carry > 0 && i < bad) bad = i;                                                \
      O[i] = o;                                                                         \
    }                                                                                   \
  } while (0)

#define CHK_MUL(ST, UT, TMIN, TMAX)                                                     \
  do {                                                                                  \
    const ST* L = (const ST*)l; const ST* R = (const ST*)r; ST* O = (ST*)out;           \
    for (int64_t i = 0; i < n; ++i) {                                                   \
      const ST a = (shape == SH_SA) ? L[0] : L[i];                                      \
      const ST b = (shape == SH_AS) ? R[0] : R[i];                                      \
      int ovf = 0;                                                                      \
      if (a > 0) { if (b > 0) { if (a > (TMAX) / b) ovf = 1; } else { if (b < (TMIN) / a) ovf = 1; } } \
      else if (b > 0) { if (a < (TMIN) / b) ovf = 1; }                                  \
      else { if (a != 0 && b < (TMAX) / a) ovf = 1; }                                   \
      if (ovf) { if (i < bad) bad = i; O[i] = 0; }                                      \
      else O[i] = (ST)((UT)a * (UT)b);                                                  \
    }                                                                                   \
  } while (0)

#define CHK_DIV(ST, IS_SIGNED, TMIN)                                                    \
  do {                                                                                  \
    const ST* L = (const ST*)l; const ST* R = (const ST*)r; ST* O = (ST*)out;           \
    for (int64_t i = 0; i < n; ++i) {                                                   \
      if (!CHK_SLOT_VALID(i)) { O[i] = 0; continue; }                                   \
      const ST a = (shape == SH_SA) ? L[0] : L[i];                                      \
      const ST b = (shape == SH_AS) ? R[0] : R[i];                                      \
      if (b == 0) { if (i < bad) bad = i; O[i] = 0; }                                   \
      else if ((IS_SIGNED) && a == (TMIN) && b == (ST)-1) O[i] = a; /* Go wraps */      \
      else O[i] = (ST)(a / b);                                                          \
    }                                                                                   \
  } while (0)

/* shiftKernelSignedImpl / UnsignedImpl K/scalar_arithmetic.go:293-379 under ScalarBinaryNotNull: an amount outside
 * [0, maxShift) leaves lhs as it is (and fails the call in the checked flavour); maxShift = bits - 1 for SIGNED types
 * (so int32 << 31 is refused), bits for unsigned.  Left shifts go through the unsigned type, right shifts are
 * arithmetic for signed operands. */
#define CHK_SHIFT(ST, UT, IS_SIGNED, LEFT, CHECKED)                                     \
  do {                                                                                  \
    const ST* L = (const ST*)l; const ST* R = (const ST*)r; ST* O = (ST*)out;           \
    const int max_shift = (int)sizeof(ST) * 8 - ((IS_SIGNED) ? 1 : 0);                  \
    for (int64_t i = 0; i < n; ++i) {                                                   \
      if (!CHK_SLOT_VALID(i)) { O[i] = 0; continue; }                                   \
      const ST a = (shape == SH_SA) ? L[0] : L[i];                                      \
      const ST b = (shape == SH_AS) ? R[0] : R[i];                                      \
      if (((IS_SIGNED) && b < 0) || (uint64_t)b >= (uint64_t)max_shift) {               \
        if ((CHECKED) && i < bad) bad = i;                                              \
        O[i] = a;                                                                       \
      } else if (LEFT) O[i] = (ST)((UT)a << (int)b);                                    \
      else O[i] = (ST)(a >> (int)b);                                                    \
    }                                                                                   \
  } while (0)

#define CHK_TYPE(ST, UT, IS_SIGNED, TMIN, TMAX)                                         \
  switch (op) {                                                                         \
    case OP_SHL: CHK_SHIFT(ST, UT, IS_SIGNED, 1, 0); break;                             \
    case OP_SHR: CHK_SHIFT(ST, UT, IS_SIGNED, 0, 0); break;                             \
    case OP_SHL_C: CHK_SHIFT(ST, UT, IS_SIGNED, 1, 1); break;                           \
    case OP_SHR_C: CHK_SHIFT(ST, UT, IS_SIGNED, 0, 1); break;                           \
    case OP_ADD_C: CHK_ADDSUB(ST, IS_SIGNED, (ST)(a & b) | ((ST)(a | b) & (ST)~o), (UT)a + (UT)b); break;      \
    case OP_SUB_C: CHK_ADDSUB(ST, IS_SIGNED, (ST)((ST)~a & b) | ((ST)~(ST)(a ^ b) & o), (UT)a - (UT)b); break; \
    case OP_MUL_C: CHK_MUL(ST, UT, TMIN, TMAX); break;                                  \
    case OP_DIV: case OP_DIV_C: CHK_DIV(ST, IS_SIGNED, TMIN); break;                    \
    default: return REF_ERR_NOT_IMPLEMENTED;                                            \
  }                                                                                     \
  break;

#define CHK_FDIV(FT)                                                                    \
  if (op != OP_DIV && op != OP_DIV_C) return REF_ERR_NOT_IMPLEMENTED;                   \
  do {                                                                                  \
    const FT* L = (const FT*)l; const FT* R = (const FT*)r; FT* O = (FT*)out;           \
    for (int64_t i = 0; i < n; ++i) {                                                   \
      if (!CHK_SLOT_VALID(i)) { O[i] = 0; continue; }                                   \
      const FT a = (shape == SH_SA) ? L[0] : L[i];                                      \
      const FT b = (shape == SH_AS) ? R[0] : R[i];                                      \
      if (op == OP_DIV_C && b == 0) { if (i < bad) bad = i; O[i] = 0; }                 \
      else O[i] = a / b;                                                                \
    }                                                                                   \
  } while (0);

int ref_arith_checked(int type, int op, int shape,
                      const void* l, const uint8_t* lvalid, int64_t loff,
                      const void* r, const uint8_t* rvalid, int64_t roff,
                      void* out, int64_t n, int64_t* first_bad) {
  int64_t bad = REF_NO_ERROR_POS;
  if (first_bad) *first_bad = bad;
  /* null scalar: "fast path if one side is entirely null" helpers.go:287,314,341 */
  if ((shape == SH_SA && l == NULL) || (shape == SH_AS && r == NULL)) return REF_OK;
  switch (type) {
    case T_I8: CHK_TYPE(int8_t, uint8_t, 1, INT8_MIN, INT8_MAX)
    case T_I16: CHK_TYPE(int16_t, uint16_t, 1, INT16_MIN, INT16_MAX)
    case T_I32: CHK_TYPE(int32_t, uint32_t, 1, INT32_MIN, INT32_MAX)
    case T_I64: CHK_TYPE(int64_t, uint64_t, 1, INT64_MIN, INT64_MAX)
    case T_U8: CHK_TYPE(uint8_t, uint8_t, 0, 0, UINT8_MAX)
    case T_U16: CHK_TYPE(uint16_t, uint16_t, 0, 0, UINT16_MAX)
    case T_U32: CHK_TYPE(uint32_t, uint32_t, 0, 0, UINT32_MAX)
    case T_U64: CHK_TYPE(uint64_t, uint64_t, 0, 0, UINT64_MAX)
    /* floating point Div / DivChecked, K/base_arithmetic.go:386-397 under ScalarBinaryNotNull: the unchecked op is
     * the IEEE quotient (x/0 = +-Inf or NaN), the checked one fails on a zero divisor (slot value 0) */
    case T_F32: CHK_FDIV(float) break;
    case T_F64: CHK_FDIV(double) break;
    default: return REF_ERR_TYPE;
  }
  if (first_bad) *first_bad = bad;
  return bad == REF_NO_ERROR_POS ? REF_OK : REF_ERR_INVALID;
}

/* AbsoluteValueChecked / NegateChecked on signed integers: K/base_arithmetic.go:295-340 under
 * ScalarUnary (every slot): v == MinInt -> errOverflow; otherwise as the unchecked kernels. */
int ref_arith_unary_checked(int type, int op, const void* in, void* out, int64_t n, int64_t* first_bad) {
  int64_t bad = REF_NO_ERROR_POS;
  if (first_bad) *first_bad = bad;
  if (op != OP_ABS_C && op != OP_NEG_C) return REF_ERR_NOT_IMPLEMENTED;
#define UC(ST, TMIN)                                                                    \
  do { const ST* I = (const ST*)in;                                                     \
       for (int64_t i = 0; i < n; ++i) if (I[i] == (TMIN) && i < bad) bad = i; } while (0)
  switch (type) {
    case T_I8: UC(int8_t, INT8_MIN); break;
    case T_I16: UC(int16_t, INT16_MIN); break;
    case T_I32: UC(int32_t, INT32_MIN); break;
    case T_I64: UC(int64_t, INT64_MIN); break;
    default: break;
  }
#undef UC
  const int rc = ref_arith_unary_same(type, op, in, out, n);
  if (rc != REF_OK) return rc;
  if (first_bad) *first_bad = bad;
  return bad == REF_NO_ERROR_POS ? REF_OK : REF_ERR_INVALID;
}

/* ====================================================================================== *
 * Integer min/max: internal/utils/min_max.go:25-210 (pure Go) = internal/utils/_lib/min_max.c:23-125.
 * Initial values (type MAX, type MIN) are what an empty input returns.
 * ====================================================================================== */
#define MINMAX_LOOP(T, TMIN, TMAX)                                                      \
  do { T lo = (TMAX), hi = (TMIN); const T* V = (const T*)in;                            \
       for (int64_t i = 0; i < n; ++i) { if (lo > V[i]) lo = V[i]; if (hi < V[i]) hi = V[i]; } \
       *(T*)min_out = lo; *(T*)max_out = hi; } while (0)
int ref_min_max(int type, const void* in, int64_t n, void* min_out, void* max_out) {
  switch (type) {
    case T_I8: MINMAX_LOOP(int8_t, INT8_MIN, INT8_MAX); break;
    case T_U8: MINMAX_LOOP(uint8_t, 0, UINT8_MAX); break;
    case T_I16: MINMAX_LOOP(int16_t, INT16_MIN, INT16_MAX); break;
    case T_U16: MINMAX_LOOP(uint16_t, 0, UINT16_MAX); break;
    case T_I32: MINMAX_LOOP(int32_t, INT32_MIN, INT32_MAX); break;
    case T_U32: MINMAX_LOOP(uint32_t, 0, UINT32_MAX); break;
    case T_I64: MINMAX_LOOP(int64_t, INT64_MIN, INT64_MAX); break;
    case T_U64: MINMAX_LOOP(uint64_t, 0, UINT64_MAX); break;
    default: return REF_ERR_TYPE;
  }
  return REF_OK;
}

/* ====================================================================================== *
 * cumulative_sum[_checked]: arrow/compute/internal/kernels/vector_cumulative.go —
 * cumulativeSumNoNulls :228-241, NoNullsChecked :243-261, WithNulls :263-290, WithNullsChecked
 * :292-324, checkedAddSigned / Unsigned :147-160, state (current, skipNulls, encounteredNull)
 * :100-106.  `state` = {current value (8 bytes, the type's own representation in the low bytes),
 * encountered_null}; it carries from chunk to chunk like the reference's kernel state.  Null and
 * dead slots are left at 0 (the reference's output comes zeroed from ctx.Allocate).  The loop
 * stops at the first overflow (checked), reporting its row.
 * ====================================================================================== */
typedef struct { uint8_t cur[8]; int64_t encountered_null; } ref_cumsum_state;
#define CUMSUM_INT(T, UT, SIGNED, TMIN, TMAX)                                                    \
  do { const T* X = (const T*)in; T* O = (T*)out; T cur; memcpy(&cur, st->cur, sizeof(T));       \
       for (int64_t i = 0; i < n; ++i) {                                                         \
         const int valid_i = !valid || bit_is_set(valid, voff + i);                              \
         if (!valid_i || st->encountered_null) {                                                 \
           if (out_valid) set_bit_to(out_valid, ooff + i, 0);                                    \
           ++nulls; O[i] = 0;                                                                    \
           if (!valid_i && !skip_nulls) st->encountered_null = 1;                                \
           continue;                                                                             \
         }                                                                                       \
         const T x = X[i];                                                                       \
         if (checked) {                                                                          \
           int ovf;                                                                              \
           if (SIGNED) ovf = (x > 0 && cur > (T)((TMAX) - x)) || (x < 0 && cur < (T)((TMIN) - x)); \
           else ovf = cur > (T)((TMAX) - x);                                                     \
           if (ovf) { bad = i; goto done; }                                                      \
         }                                                                                       \
         cur = (T)((UT)cur + (UT)x);                                                             \
         O[i] = cur;                                                                             \
         if (out_valid) set_bit_to(out_valid, ooff + i, 1);                                      \
       }                                                                                         \
       memcpy(st->cur, &cur, sizeof(T)); } while (0)
#define CUMSUM_FLT(T)                                                                            \
  do { const T* X = (const T*)in; T* O = (T*)out; T cur; memcpy(&cur, st->cur, sizeof(T));       \
       for (int64_t i = 0; i < n; ++i) {                                                         \
         const int valid_i = !valid || bit_is_set(valid, voff + i);                              \
         if (!valid_i || st->encountered_null) {                                                 \
           if (out_valid) set_bit_to(out_valid, ooff + i, 0);                                    \
           ++nulls; O[i] = 0;                                                                    \
           if (!valid_i && !skip_nulls) st->encountered_null = 1;                                \
           continue;                                                                             \
         }                                                                                       \
         cur = cur + X[i];                                                                       \
         O[i] = cur;                                                                             \
         if (out_valid) set_bit_to(out_valid, ooff + i, 1);                                      \
       }                                                                                         \
       memcpy(st->cur, &cur, sizeof(T)); } while (0)
int ref_cumulative_sum(int type, const void* in, const uint8_t* valid, int64_t voff, int64_t n, int skip_nulls, int checked,
                       void* out, uint8_t* out_valid, int64_t ooff, void* state, int64_t* null_count, int64_t* first_bad) {
  ref_cumsum_state* st = (ref_cumsum_state*)state;
  int64_t nulls = 0, bad = REF_NO_ERROR_POS;
  switch (type) {
    case T_I8: CUMSUM_INT(int8_t, uint8_t, 1, INT8_MIN, INT8_MAX); break;
    case T_U8: CUMSUM_INT(uint8_t, uint8_t, 0, 0, UINT8_MAX); break;
    case T_I16: CUMSUM_INT(int16_t, uint16_t, 1, INT16_MIN, INT16_MAX); break;
    case T_U16: CUMSUM_INT(uint16_t, uint16_t, 0, 0, UINT16_MAX); break;
    case T_I32: CUMSUM_INT(int32_t, uint32_t, 1, INT32_MIN, INT32_MAX); break;
    case T_U32: CUMSUM_INT(uint32_t, uint32_t, 0, 0, UINT32_MAX); break;
    case T_I64: CUMSUM_INT(int64_t, uint64_t, 1, INT64_MIN, INT64_MAX); break;
    case T_U64: CUMSUM_INT(uint64_t, uint64_t, 0, 0, UINT64_MAX); break;
    case T_F32: CUMSUM_FLT(float); break;
    case T_F64: CUMSUM_FLT(double); break;
    default: return REF_ERR_TYPE;
  }
done:
  if (null_count) *null_count = nulls;
  if (first_bad) *first_bad = bad;
  return bad == REF_NO_ERROR_POS ? REF_OK : REF_ERR_INVALID;
}

/* ====================================================================================== *
 * Numeric casts: the loop K/cast_numeric.go:101-131 (= K/_lib/cast_numeric.cc:22-101) framed by
 * the safe-cast checks of K/numeric_cast.go:37-71 —
 *   int -> int    intsCanFit (K/helpers.go:545-578) bounds from getSafeMinMax* (:496-543),
 *                 intsInRange (:580-653) incl. its "whole input range fits" early return;
 *   int -> float  checkIntToFloatTrunc (K/numeric_cast.go:698-729);
 *   float -> int  checkFloatTrunc / wasTrunc (K/numeric_cast.go:587-660), evaluated AFTER the loop.
 * Every slot is converted; only valid slots are checked; *first_bad = lowest failing row.
 * float -> int of an unrepresentable value is undefined in C and varies between the reference's
 * own AVX2 / SSE4 / scalar-tail code; this restates what Go's conversion does on amd64
 * (CVTTSD2SQ "integer indefinite" + wrap), which is also what the CUDA path documents.
 * ====================================================================================== */
static int64_t cvtt64(double v) {
  if (!(v >= -9223372036854775808.0 && v < 9223372036854775808.0)) return INT64_MIN;
  return (int64_t)v;
}
static uint64_t f2u64(double v) {
  if (v >= 9223372036854775808.0) return (uint64_t)cvtt64(v - 9223372036854775808.0) ^ 0x8000000000000000ull;
  return (uint64_t)cvtt64(v);
}
static int cast_is_int(int t) { return t >= T_U8 && t <= T_I64; }
static int cast_is_signed(int t) { return t == T_I8 || t == T_I16 || t == T_I32 || t == T_I64; }
static int cast_width(int t) {
  switch (t) { case T_U8: case T_I8: return 1; case T_U16: case T_I16: return 2;
               case T_U32: case T_I32: case T_F32: return 4; case T_U64: case T_I64: case T_F64: return 8; default: return 0; }
}
typedef __int128 i128;
static i128 int_min_of(int t) { return cast_is_signed(t) ? -((i128)1 << (8 * cast_width(t) - 1)) : 0; }
static i128 int_max_of(int t) { return cast_is_signed(t) ? ((i128)1 << (8 * cast_width(t) - 1)) - 1 : ((i128)1 << (8 * cast_width(t))) - 1; }

/* carriers: ints as i128 (exact), floats as double (float32 -> double is exact) */
static i128 load_int(int t, const void* p, int64_t i) {
  switch (t) {
    case T_U8: return ((const uint8_t*)p)[i];   case T_I8: return ((const int8_t*)p)[i];
    case T_U16: return ((const uint16_t*)p)[i]; case T_I16: return ((const int16_t*)p)[i];
    case T_U32: return ((const uint32_t*)p)[i]; case T_I32: return ((const int32_t*)p)[i];
    case T_U64: return ((const uint64_t*)p)[i]; default: return ((const int64_t*)p)[i];
  }
}
static void store_int_bits(int t, void* p, int64_t i, uint64_t bits) {
  switch (cast_width(t)) {
    case 1: ((uint8_t*)p)[i] = (uint8_t)bits; break;
    case 2: ((uint16_t*)p)[i] = (uint16_t)bits; break;
    case 4: ((uint32_t*)p)[i] = (uint32_t)bits; break;
    default: ((uint64_t*)p)[i] = bits; break;
  }
}

int ref_cast_numeric(int itype, int otype, const void* in, const uint8_t* valid, int64_t voff, void* out, int64_t n,
                     int allow_int_overflow, int allow_float_truncate, int64_t* first_bad) {
  int64_t bad = REF_NO_ERROR_POS;
  if (first_bad) *first_bad = bad;
  const int wi = cast_width(itype), wo = cast_width(otype);
  if (wi == 0 || wo == 0) return REF_ERR_TYPE;
  if (n < 0) return REF_ERR_INVALID;
  if (itype == otype) { memcpy(out, in, (size_t)n * wi); return REF_OK; }
  const int iint = cast_is_int(itype), oint = cast_is_int(otype);
  /* bounds for the integer-input checks */
  int check_range = 0;
  i128 lo = 0, hi = 0;
  if (iint && oint && !allow_int_overflow) {
    lo = int_min_of(itype); hi = int_max_of(itype);
    if (cast_is_signed(itype)) {
      if (cast_is_signed(otype)) { if (wi > wo) { lo = int_min_of(otype); hi = int_max_of(otype); } }
      else { lo = 0; if (wi > wo) hi = int_max_of(otype); }
    } else {
      if (cast_is_signed(otype)) { if (wi >= wo) hi = int_max_of(otype); }
      else { if (wi > wo) hi = int_max_of(otype); }
    }
    check_range = !(lo <= int_min_of(itype) && hi >= int_max_of(itype));
  } else if (iint && !oint && !allow_float_truncate) {
    const int mant = (otype == T_F32) ? 24 : 53;
    if (8 * wi > mant && !(wi == 4 && otype == T_F64)) {  /* int8/16 never; int32/uint32 only -> float32 */
      hi = (i128)1 << mant; lo = cast_is_signed(itype) ? -hi : 0; check_range = 1;
    }
  }
  const int check_trunc = !iint && oint && !allow_float_truncate;
  for (int64_t i = 0; i < n; ++i) {
    const int is_valid = !valid || ((valid[(voff + i) >> 3] >> ((voff + i) & 7)) & 1);
    if (iint) {
      const i128 v = load_int(itype, in, i);
      if (check_range && is_valid && (v < lo || v > hi) && i < bad) bad = i;
      if (oint) store_int_bits(otype, out, i, (uint64_t)v);
      else if (otype == T_F32) ((float*)out)[i] = cast_is_signed(itype) ? (float)(int64_t)v : (float)(uint64_t)v;
      else ((double*)out)[i] = cast_is_signed(itype) ? (double)(int64_t)v : (double)(uint64_t)v;
    } else {
      const double d = (itype == T_F32) ? (double)((const float*)in)[i] : ((const double*)in)[i];
      if (!oint) {
        if (otype == T_F32) ((float*)out)[i] = (float)d; else ((double*)out)[i] = d;
      } else {
        const uint64_t bits = (otype == T_U64) ? f2u64(d) : (uint64_t)cvtt64(d);
        store_int_bits(otype, out, i, bits);
        if (check_trunc && is_valid) {
          const i128 o = load_int(otype, out, i);   /* the stored (wrapped) OutT value */
          int trunc;
          if (itype == T_F32) {
            const float back = cast_is_signed(otype) ? (float)(int64_t)o : (float)(uint64_t)o;
            trunc = !(back == ((const float*)in)[i]);
          } else {
            const double back = cast_is_signed(otype) ? (double)(int64_t)o : (double)(uint64_t)o;
            trunc = !(back == d);
          }
          if (trunc && i < bad) bad = i;
        }
      }
    }
  }
  if (first_bad) *first_bad = bad;
  return bad == REF_NO_ERROR_POS ? REF_OK : REF_ERR_INVALID;
}

/* ====================================================================================== *
 * Comparisons: K/_lib/scalar_comparison.cc:63-206 (prefix bits up to the next byte boundary,
 * 32-wide batches packed LSB-first, tail bits), driver K/scalar_comparisons.go:199-218,
 * LT/LE by flipping K/../scalar_compare.go:73-99.  Bits outside [offset, offset+n) keep
 * their value.
 * ====================================================================================== */
#define CMP_LOOP(T)                                                                     \
  do {                                                                                  \
    const T* L = (const T*)l; const T* R = (const T*)r;                                 \
    for (int64_t i = 0; i < n; ++i) {                                                   \
      const T a = (shape == SH_SA) ? L[0] : L[i];                                       \
      const T b = (shape == SH_AS) ? R[0] : R[i];                                       \
      int v;                                                                            \
      switch (cmp) {                                                                    \
        case CMP_EQ: v = (a == b); break;                                               \
        case CMP_NE: v = (a != b); break;                                               \
        case CMP_GT: v = (a > b); break;                                                \
        case CMP_GE: v = (a >= b); break;                                               \
        case CMP_LT: v = (b > a); break;                                                \
        default: v = (b >= a); break;                                                   \
      }                                                                                 \
      set_bit_to(out, (int64_t)(offset % 8) + i, v);                                    \
    }                                                                                   \
  } while (0)

int ref_compare(int type, int cmp, int shape, const void* l, const void* r, uint8_t* out, int64_t n, int offset) {
  if (cmp < 0 || cmp > CMP_LE) return REF_ERR_INVALID;
  switch (type) {
    case T_U8: CMP_LOOP(uint8_t); break;
    case T_I8: CMP_LOOP(int8_t); break;
    case T_U16: CMP_LOOP(uint16_t); break;
    case T_I16: CMP_LOOP(int16_t); break;
    case T_U32: CMP_LOOP(uint32_t); break;
    case T_I32: CMP_LOOP(int32_t); break;
    case T_U64: CMP_LOOP(uint64_t); break;
    case T_I64: CMP_LOOP(int64_t); break;
    case T_F32: CMP_LOOP(float); break;
    case T_F64: CMP_LOOP(double); break;
    default: return REF_ERR_TYPE;
  }
  return REF_OK;
}

/* ====================================================================================== *
 * Bitmaps: arrow/bitutil/bitmaps.go:527-639 (BitmapAnd/Or/Xor/AndNot/Xnor, any offsets,
 * aligned path -> _lib/bitmap_ops.c:24-46), CopyBitmap/InvertBitmap :483-491,
 * SetBitsTo bitutil.go:158, CountSetBits :89.
 * ====================================================================================== */
int ref_bitmap_op(int bitop, const uint8_t* l, int64_t loff, const uint8_t* r, int64_t roff,
                  uint8_t* out, int64_t ooff, int64_t n) {
  for (int64_t i = 0; i < n; ++i) {
    const int a = bit_is_set(l, loff + i), b = bit_is_set(r, roff + i);
    int v;
    switch (bitop) {
      case 0: v = a & b; break;
      case 1: v = a | b; break;
      case 2: v = a ^ b; break;
      case 3: v = a & !b; break;
      case 4: v = !(a ^ b); break;
      default: return REF_ERR_INVALID;
    }
    set_bit_to(out, ooff + i, v);
  }
  return REF_OK;
}
void ref_bitmap_copy(const uint8_t* src, int64_t soff, int64_t n, uint8_t* dst, int64_t doff) {
  for (int64_t i = 0; i < n; ++i) set_bit_to(dst, doff + i, bit_is_set(src, soff + i));
}
void ref_bitmap_invert(const uint8_t* src, int64_t soff, int64_t n, uint8_t* dst, int64_t doff) {
  for (int64_t i = 0; i < n; ++i) set_bit_to(dst, doff + i, !bit_is_set(src, soff + i));
}
void ref_bitmap_set(uint8_t* bits, int64_t off, int64_t n, int value) {
  for (int64_t i = 0; i < n; ++i) set_bit_to(bits, off + i, value);
}
int64_t ref_bitmap_popcount(const uint8_t* bits, int64_t off, int64_t n) {
  int64_t c = 0;
  int64_t i = 0;
  /* byte-at-a-time once aligned: same answer as the bit loop, fast enough for 100M-bit masks */
  while (i < n && ((off + i) & 7)) { c += bit_is_set(bits, off + i); ++i; }
  while (i + 8 <= n) { c += __builtin_popcount(bits[(off + i) >> 3]); i += 8; }
  while (i < n) { c += bit_is_set(bits, off + i); ++i; }
  return c;
}

/* Kleene: K/scalar_boolean.go:29-65 computeKleene; word lambdas :103-105 (and),
 * :180-182 (or), :290-292 (and_not). */
int ref_kleene(int kop, const uint8_t* lvalid, const uint8_t* ldata, int64_t loff,
               const uint8_t* rvalid, const uint8_t* rdata, int64_t roff,
               uint8_t* out_valid, uint8_t* out_data, int64_t ooff, int64_t n) {
  for (int64_t i = 0; i < n; ++i) {
    const int lv = lvalid ? bit_is_set(lvalid, loff + i) : 1, ld = bit_is_set(ldata, loff + i);
    const int rv = rvalid ? bit_is_set(rvalid, roff + i) : 1, rd = bit_is_set(rdata, roff + i);
    const int lt = lv & ld, lf = lv & !ld, rt = rv & rd, rf = rv & !rd;
    int ov, od;
    switch (kop) {
      case 0: ov = lf | rf | (lt & rt); od = lt & rt; break;
      case 1: ov = lt | rt | (lf & rf); od = lt | rt; break;
      case 2: ov = lf | rt | (lt & rf); od = lt & rf; break;
      default: return REF_ERR_INVALID;
    }
    set_bit_to(out_valid, ooff + i, ov);
    set_bit_to(out_data, ooff + i, od);
  }
  return REF_OK;
}

/* ====================================================================================== *
 * Filter: K/vector_selection.go
 *   getFilterOutputSize :57-81 — popcount(mask ∧ valid) for DropNulls, popcount(mask ∨ ¬valid)
 *     for EmitNulls, CountSetBits(mask) without mask validity.
 *   primitiveFilterImpl :267-395 walks ≤64-row blocks with four block-level fast paths
 *     (:303-320) that are shortcuts of the per-row rule of the default branch (:321-392):
 *       selected  (maskValid ∧ mask)        -> WriteValue; validity = 1 or valuesValid[i]
 *       EmitNulls ∧ ¬maskValid              -> WriteNull (value 0 :417-421, validity 0)
 *       otherwise                            -> skipped
 *     and the no-null path :275-283 copies set-bit runs.  All block counters advance in lock
 *     step with length min(64, remaining) (internal/bitutils/bit_block_counter.go:82-92,
 *     144-166, 212-224), so the row rule below is the same function.  We keep the block
 *     structure anyway so that the block-level branches are exercised as written.
 *   PrimitiveFilter :449-520 — validity buffer only when either input may have nulls; values
 *     under selected-but-null value slots are copied as they are.
 * ====================================================================================== */
int64_t ref_filter_output_size(const uint8_t* mask, const uint8_t* mvalid, int64_t moff, int64_t n, int null_selection) {
  int64_t c = 0;
  for (int64_t i = 0; i < n; ++i) {
    const int m = bit_is_set(mask, moff + i);
    const int v = mvalid ? bit_is_set(mvalid, moff + i) : 1;
    c += null_selection ? (m | !v) : (m & v);
  }
  return c;
}

static inline void copy_elem(void* out, int64_t opos, const void* vals, int64_t ipos, int w) {
  memcpy((char*)out + opos * w, (const char*)vals + ipos * w, (size_t)w);
}

int ref_filter_primitive(int bit_width, const void* vals, const uint8_t* vvalid, int64_t voff,
                         const uint8_t* mask, const uint8_t* mvalid, int64_t moff, int64_t n,
                         int null_selection, void* out, uint8_t* out_valid, int64_t* out_len, int64_t* out_nulls) {
  if (bit_width == 1) {
    /* boolean values (boolFilterWriter :423-447).  Row rule of the default branch; NB the
     * reference's WriteValue does not advance its output position (:433-436) — we restate the
     * documented behaviour (what every other width does), see DESIGN.md "reference quirks". */
    int64_t opos = 0, nulls = 0;
    for (int64_t i = 0; i < n; ++i) {
      const int mv = mvalid ? bit_is_set(mvalid, moff + i) : 1;
      const int m = bit_is_set(mask, moff + i);
      if (mv && m) {
        const int v = vvalid ? bit_is_set(vvalid, voff + i) : 1;
        if (out_valid) set_bit_to(out_valid, opos, v);
        nulls += !v;
        set_bit_to((uint8_t*)out, opos, bit_is_set((const uint8_t*)vals, voff + i));
        ++opos;
      } else if (!mv && null_selection == 1) {
        if (out_valid) set_bit_to(out_valid, opos, 0);
        set_bit_to((uint8_t*)out, opos, 0);
        ++nulls;
        ++opos;
      }
    }
    if (out_len) *out_len = opos;
    if (out_nulls) *out_nulls = nulls;
    return REF_OK;
  }
  if (bit_width != 8 && bit_width != 16 && bit_width != 32 && bit_width != 64) return REF_ERR_TYPE;
  const int w = bit_width / 8;
  int64_t opos = 0, nulls = 0;
  const char* vbase = (const char*)vals + voff * w;
  for (int64_t blk = 0; blk < n; blk += 64) {
    const int64_t len = (n - blk < 64) ? (n - blk) : 64;
    int64_t filter_cnt = 0, data_cnt = 0;
    for (int64_t i = blk; i < blk + len; ++i) {
      filter_cnt += bit_is_set(mask, moff + i) & (mvalid ? bit_is_set(mvalid, moff + i) : 1);
      data_cnt += vvalid ? bit_is_set(vvalid, voff + i) : 1;
    }
    if (filter_cnt == len && data_cnt == len) {                       /* :303-308 */
      if (out_valid) for (int64_t k = 0; k < len; ++k) set_bit_to(out_valid, opos + k, 1);
      memcpy((char*)out + opos * w, vbase + blk * w, (size_t)(len * w));
      opos += len;
    } else if (filter_cnt == len) {                                   /* :309-316 */
      for (int64_t k = 0; k < len; ++k) {
        const int v = bit_is_set(vvalid, voff + blk + k);
        set_bit_to(out_valid, opos + k, v);
        nulls += !v;
      }
      memcpy((char*)out + opos * w, vbase + blk * w, (size_t)(len * w));
      opos += len;
    } else if (filter_cnt == 0 && null_selection == 0) {              /* :317-320 */
      continue;
    } else {                                                          /* :321-392 */
      for (int64_t i = blk; i < blk + len; ++i) {
        const int mv = mvalid ? bit_is_set(mvalid, moff + i) : 1;
        const int m = bit_is_set(mask, moff + i);
        if (mv && m) {
          const int v = vvalid ? bit_is_set(vvalid, voff + i) : 1;
          if (out_valid) set_bit_to(out_valid, opos, v);
          nulls += !v;
          copy_elem(out, opos, vbase, i, w);
          ++opos;
        } else if (!mv && null_selection == 1) {
          if (out_valid) set_bit_to(out_valid, opos, 0);
          memset((char*)out + opos * w, 0, (size_t)w);
          ++nulls;
          ++opos;
        }
      }
    }
  }
  if (out_len) *out_len = opos;
  if (out_nulls) *out_nulls = nulls;
  return REF_OK;
}

/* GetTakeIndices: K/vector_selection.go:102-236.  Per row: valid∧true -> index; EmitNulls∧¬valid
 * -> null slot (builder leaves the value 0); DropNulls drops nulls.  NB: the reference's
 * EmitNulls loop only advances its validity block counter on non-empty blocks (:131-139), so
 * after skipping an all-false block its all-valid shortcut can look at a stale block; we
 * restate the documented row rule (the comment at :111-115), which is what every other branch
 * implements.  See DESIGN.md "reference quirks". */
int ref_take_indices(int index_width, const uint8_t* mask, const uint8_t* mvalid, int64_t moff, int64_t n,
                     int null_selection, void* out_idx, uint8_t* out_valid, int64_t* out_len) {
  if (index_width != 16 && index_width != 32) return REF_ERR_TYPE;
  int64_t opos = 0;
  for (int64_t i = 0; i < n; ++i) {
    const int mv = mvalid ? bit_is_set(mvalid, moff + i) : 1;
    const int m = bit_is_set(mask, moff + i);
    if (mv && m) {
      if (index_width == 16) ((uint16_t*)out_idx)[opos] = (uint16_t)i; else ((uint32_t*)out_idx)[opos] = (uint32_t)i;
      if (out_valid) set_bit_to(out_valid, opos, 1);
      ++opos;
    } else if (!mv && null_selection == 1) {
      if (index_width == 16) ((uint16_t*)out_idx)[opos] = 0; else ((uint32_t*)out_idx)[opos] = 0;
      if (out_valid) set_bit_to(out_valid, opos, 0);
      ++opos;
    }
  }
  if (out_len) *out_len = opos;
  return REF_OK;
}

/* ====================================================================================== *
 * Take: K/vector_selection.go:1162-1192 PrimitiveTake
 *   checkIndexBounds K/helpers.go:929-981 — valid index slots only; signed: idx < 0 or
 *     idx >= len; error names the first offender in row order.
 *   primitiveTakeImpl :878-988 — out[i] = values[idx[i]] for slots that are valid in the index
 *     and (if values have nulls) in values; other slots are left as allocated (zero).
 *     Indices are reinterpreted as unsigned of the same width (:1147-1159).
 * ====================================================================================== */
static inline uint64_t load_index(const void* idx, int64_t i, int width, int is_signed, int* negative) {
  *negative = 0;
  switch (width) {
    case 8: { if (is_signed) { int8_t v = ((const int8_t*)idx)[i]; *negative = v < 0; return (uint64_t)(int64_t)v; } return ((const uint8_t*)idx)[i]; }
    case 16: { if (is_signed) { int16_t v = ((const int16_t*)idx)[i]; *negative = v < 0; return (uint64_t)(int64_t)v; } return ((const uint16_t*)idx)[i]; }
    case 32: { if (is_signed) { int32_t v = ((const int32_t*)idx)[i]; *negative = v < 0; return (uint64_t)(int64_t)v; } return ((const uint32_t*)idx)[i]; }
    default: { if (is_signed) { int64_t v = ((const int64_t*)idx)[i]; *negative = v < 0; return (uint64_t)v; } return ((const uint64_t*)idx)[i]; }
  }
}

int ref_take_primitive(int bit_width, const void* vals, const uint8_t* vvalid, int64_t voff, int64_t vlen,
                       int idx_width, int idx_signed, const void* idx, const uint8_t* ivalid, int64_t ioff,
                       int64_t n, int bounds_check, void* out, uint8_t* out_valid,
                       int64_t* out_nulls, int64_t* bad_pos, int64_t* bad_index) {
  if (bit_width != 1 && bit_width != 8 && bit_width != 16 && bit_width != 32 && bit_width != 64) return REF_ERR_TYPE;
  if (idx_width != 8 && idx_width != 16 && idx_width != 32 && idx_width != 64) return REF_ERR_INDEX;
  const int w = bit_width / 8; /* 0 for boolean values: booleanTakeImpl :990-1074 */
  if (bad_pos) *bad_pos = REF_NO_ERROR_POS;
  if (bounds_check) {
    for (int64_t i = 0; i < n; ++i) {
      if (ivalid && !bit_is_set(ivalid, ioff + i)) continue;
      int neg;
      const uint64_t v = load_index(idx, i, idx_width, idx_signed, &neg);
      if (neg || v >= (uint64_t)vlen) {
        if (bad_pos) *bad_pos = i;
        if (bad_index) *bad_index = (int64_t)v;
        return REF_ERR_INDEX;
      }
    }
  }
  int64_t nulls = 0;
  const char* vbase = (const char*)vals + voff * w;
  for (int64_t i = 0; i < n; ++i) {
    int neg;
    const int iv = ivalid ? bit_is_set(ivalid, ioff + i) : 1;
    int ok = iv;
    uint64_t v = 0;
    if (iv) {
      v = load_index(idx, i, idx_width, idx_signed, &neg);
      if (idx_width < 64) v &= ((1ull << idx_width) - 1);   /* unsigned reinterpretation */
      if (vvalid) ok = bit_is_set(vvalid, voff + (int64_t)v);
    }
    if (ok) {
      if (bit_width == 1) set_bit_to((uint8_t*)out, i, bit_is_set((const uint8_t*)vals, voff + (int64_t)v));
      else copy_elem(out, i, vbase, (int64_t)v, w);
      if (out_valid) set_bit_to(out_valid, i, 1);
    } else {
      if (bit_width == 1) set_bit_to((uint8_t*)out, i, 0);
      else memset((char*)out + i * w, 0, (size_t)w);
      if (out_valid) set_bit_to(out_valid, i, 0);
      ++nulls;
    }
  }
  if (out_nulls) *out_nulls = nulls;
  return REF_OK;
}

/* ====================================================================================== *
 * Parity helpers (definitions shared with arrow_go_b200/csrc/util.cu)
 * ====================================================================================== */
static inline uint64_t mix64(uint64_t z) {
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
uint64_t ref_checksum64(const void* buf, size_t n_words) {
  const uint64_t* p = (const uint64_t*)buf;
  uint64_t acc = 0;
  for (size_t i = 0; i < n_words; ++i) acc += mix64((uint64_t)i) * p[i];
  return acc;
}
void ref_generate(int kind, uint64_t seed, int64_t lo, int64_t hi, void* out, size_t n) {
  const uint64_t span = (uint64_t)(hi - lo) + 1;
  for (size_t i = 0; i < n; ++i) {
    const uint64_t z = mix64(seed + (uint64_t)i);
    const uint64_t k = ((z >> 32) * span) >> 32; /* uniform in [0, span) for span <= 2^32 */
    switch (kind) {
      case 0: ((uint64_t*)out)[i] = z; break;
      case 1: ((int64_t*)out)[i] = lo + (int64_t)k; break;
      case 2: ((int32_t*)out)[i] = (int32_t)(lo + (int64_t)k); break;
      case 3: ((double*)out)[i] = (double)(lo + (int64_t)k); break;
      case 4: {
        /* bit i set with probability lo/hi */
        const int bit = ((z >> 32) * (uint64_t)hi >> 32) < (uint64_t)lo;
        uint8_t* b = (uint8_t*)out;
        if ((i & 7) == 0) b[i >> 3] = 0;
        b[i >> 3] |= (uint8_t)(bit << (i & 7));
        break;
      }
      case 5: ((int32_t*)out)[i] = (int32_t)(lo + (int64_t)(((unsigned __int128)i * span) / n)); break;
      case 6: ((int32_t*)out)[i] = (int32_t)(lo + (int64_t)(((unsigned __int128)(n - 1 - i) * span) / n)); break;
      default: return;
    }
  }
}


/* ------------------------------------------------------------------------------------------------------------
 * sort_indices, single fixed-width column (kernels.SortIndices for one key: arrow/compute/internal/kernels/
 * vector_sort.go:385-481; arraySortOneColumnRange vector_sort_internal.go:250-300 = partitionNullsOnly :36-87 +
 * partitionNullLikes :89-150 + slices.SortStableFunc over the finite range with compareRowsForKey
 * (vector_sort_physical.go: compareKeyedNulls / compareFloatNaNs / compareOrdered, vector_sort_support.go:100-170)).
 *   NullsAtEnd  : [finite in key order | NaN in row order | null in row order]
 *   NullsAtStart: [null in row order | NaN in row order | finite in key order]
 * Stable; Descending reverses the key order only (ties stay in row order); -0.0 == +0.0.
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct { int type; const void* vals; int64_t voff; int desc; } sort_ctx_t;

static int sort_is_nan(const sort_ctx_t* c, uint64_t i) {
  if (c->type == T_F32) { float v = ((const float*)c->vals)[c->voff + (int64_t)i]; return v != v; }
  if (c->type == T_F64) { double v = ((const double*)c->vals)[c->voff + (int64_t)i]; return v != v; }
  return 0;
}

/* compareOrdered on two finite rows */
static int sort_cmp(const sort_ctx_t* c, uint64_t a, uint64_t b) {
  int r = 0;
#define CMP_AS(T) { const T x = ((const T*)c->vals)[c->voff + (int64_t)a], y = ((const T*)c->vals)[c->voff + (int64_t)b]; r = x < y ? -1 : (x > y ? 1 : 0); }
  switch (c->type) {
    case T_I8: CMP_AS(int8_t) break;   case T_U8: CMP_AS(uint8_t) break;
    case T_I16: CMP_AS(int16_t) break; case T_U16: CMP_AS(uint16_t) break;
    case T_I32: CMP_AS(int32_t) break; case T_U32: CMP_AS(uint32_t) break;
    case T_I64: CMP_AS(int64_t) break; case T_U64: CMP_AS(uint64_t) break;
    case T_F32: CMP_AS(float) break;   case T_F64: CMP_AS(double) break;
    default: break;
  }
#undef CMP_AS
  return c->desc ? -r : r;
}

/* stable merge sort of idx[lo,hi) (slices.SortStableFunc) */
static void sort_merge(const sort_ctx_t* c, uint64_t* idx, uint64_t* tmp, int64_t lo, int64_t hi) {
  if (hi - lo < 2) return;
  if (hi - lo <= 16) { /* insertion sort: stable */
    for (int64_t i = lo + 1; i < hi; ++i) {
      const uint64_t v = idx[i];
      int64_t j = i;
      while (j > lo && sort_cmp(c, idx[j - 1], v) > 0) { idx[j] = idx[j - 1]; --j; }
      idx[j] = v;
    }
    return;
  }
  const int64_t mid = lo + (hi - lo) / 2;
  sort_merge(c, idx, tmp, lo, mid);
  sort_merge(c, idx, tmp, mid, hi);
  int64_t i = lo, j = mid, k = lo;
  while (i < mid && j < hi) tmp[k++] = (sort_cmp(c, idx[j], idx[i]) < 0) ? idx[j++] : idx[i++];  /* ties take the left run */
  while (i < mid) tmp[k++] = idx[i++];
  while (j < hi) tmp[k++] = idx[j++];
  for (k = lo; k < hi; ++k) idx[k] = tmp[k];
}

int ref_sort_indices(int type, const void* vals, const uint8_t* valid, int64_t voff, int64_t n, int order, int null_placement, uint64_t* out,
                     int64_t* null_count, int64_t* nan_count) {
  if (n < 0 || voff < 0 || (order != 0 && order != 1) || (null_placement != 0 && null_placement != 1)) return REF_ERR_INVALID;
  sort_ctx_t c = {type, vals, voff, order};
  int64_t n_null = 0, n_nan = 0;
  for (int64_t i = 0; i < n; ++i) {
    if (valid && !bit_is_set(valid, voff + i)) ++n_null;
    else if (sort_is_nan(&c, (uint64_t)i)) ++n_nan;
  }
  const int64_t n_fin = n - n_null - n_nan;
  int64_t pf = null_placement ? n_null + n_nan : 0, pn = null_placement ? n_null : n_fin, pz = null_placement ? 0 : n_fin + n_nan;
  const int64_t fin_lo = pf;
  for (int64_t i = 0; i < n; ++i) {
    if (valid && !bit_is_set(valid, voff + i)) out[pz++] = (uint64_t)i;
    else if (sort_is_nan(&c, (uint64_t)i)) out[pn++] = (uint64_t)i;
    else out[pf++] = (uint64_t)i;
  }
  uint64_t* tmp = (uint64_t*)malloc((size_t)(n > 0 ? n : 1) * sizeof(uint64_t));
  if (!tmp) return REF_ERR_INVALID;
  sort_merge(&c, out, tmp, fin_lo, fin_lo + n_fin);
  free(tmp);
  if (null_count) *null_count = n_null;
  if (nan_count) *nan_count = n_nan;
  return REF_OK;
}


/* ------------------------------------------------------------------------------------------------------------
 * is_in / unique for fixed-width values, keyed by raw bytes (kernels/scalar_set_lookup.go:112-413: SetLookupState
 * over uint8/16/32/64 memo tables, isInKernelExec :373-413; kernels/vector_hash.go: unique = distinct values in
 * order of first appearance, null kept once).  Restated with a sort-free quadratic-safe approach: an open hash of
 * the raw 64-bit keys (test sizes are small).
 * ------------------------------------------------------------------------------------------------------------ */
static uint64_t raw_key_at(int bit_width, const void* vals, int64_t i) {
  switch (bit_width) {
    case 8: return ((const uint8_t*)vals)[i];
    case 16: return ((const uint16_t*)vals)[i];
    case 32: return ((const uint32_t*)vals)[i];
    default: return ((const uint64_t*)vals)[i];
  }
}
typedef struct { uint64_t* keys; int64_t* first; uint8_t* used; uint64_t mask; } rset_t;
static int rset_init(rset_t* s, int64_t entries) {
  uint64_t slots = 16;
  while (slots < (uint64_t)entries * 2 + 2) slots <<= 1;
  s->keys = (uint64_t*)calloc(slots, 8); s->first = (int64_t*)calloc(slots, 8); s->used = (uint8_t*)calloc(slots, 1); s->mask = slots - 1;
  return s->keys && s->first && s->used;
}
static void rset_free(rset_t* s) { free(s->keys); free(s->first); free(s->used); }
static int64_t rset_find(const rset_t* s, uint64_t key) { /* first row or -1 */
  uint64_t h = mix64(key) & s->mask;
  while (s->used[h]) { if (s->keys[h] == key) return s->first[h]; h = (h + 1) & s->mask; }
  return -1;
}
static void rset_add(rset_t* s, uint64_t key, int64_t row) {
  uint64_t h = mix64(key) & s->mask;
  while (s->used[h]) { if (s->keys[h] == key) return; h = (h + 1) & s->mask; }
  s->used[h] = 1; s->keys[h] = key; s->first[h] = row;
}

int ref_is_in(int bit_width, const void* vals, const uint8_t* valid, int64_t off, int64_t n, const void* set, const uint8_t* set_valid,
              int64_t set_off, int64_t set_n, int null_behavior, uint8_t* out_data, uint8_t* out_valid, int64_t* out_nulls) {
  if (bit_width != 8 && bit_width != 16 && bit_width != 32 && bit_width != 64) return REF_ERR_NOT_IMPLEMENTED;
  rset_t s;
  if (!rset_init(&s, set_n)) return REF_ERR_INVALID;
  int has_null = 0;
  for (int64_t i = 0; i < set_n; ++i) {
    if (set_valid && !bit_is_set(set_valid, set_off + i)) has_null = 1;
    else rset_add(&s, raw_key_at(bit_width, set, set_off + i), i);
  }
  int64_t nulls = 0;
  for (int64_t i = 0; i < n; ++i) {
    int d = 0, v = 1;
    if (valid && !bit_is_set(valid, off + i)) {
      if (null_behavior == 0 && has_null) d = 1;                                        /* MATCH */
      else if (null_behavior == 1 || (null_behavior == 0 && !has_null)) d = 0;          /* SKIP */
      else v = 0;
    } else if (rset_find(&s, raw_key_at(bit_width, vals, off + i)) >= 0) {
      d = 1;
    } else if (null_behavior == 3 && has_null) {                                        /* INCONCLUSIVE */
      v = 0;
    }
    if (d) out_data[i >> 3] |= (uint8_t)(1u << (i & 7)); else out_data[i >> 3] &= (uint8_t)~(1u << (i & 7));
    if (out_valid) { if (v) out_valid[i >> 3] |= (uint8_t)(1u << (i & 7)); else out_valid[i >> 3] &= (uint8_t)~(1u << (i & 7)); }
    nulls += !v;
  }
  rset_free(&s);
  if (out_nulls) *out_nulls = nulls;
  return REF_OK;
}

int ref_unique(int bit_width, const void* vals, const uint8_t* valid, int64_t off, int64_t n, void* out, uint8_t* out_valid, int64_t* out_len,
               int64_t* out_nulls) {
  if (bit_width != 8 && bit_width != 16 && bit_width != 32 && bit_width != 64) return REF_ERR_NOT_IMPLEMENTED;
  rset_t s;
  if (!rset_init(&s, n)) return REF_ERR_INVALID;
  const int w = bit_width / 8;
  int64_t k = 0, nulls = 0;
  int seen_null = 0;
  for (int64_t i = 0; i < n; ++i) {
    if (valid && !bit_is_set(valid, off + i)) {
      if (seen_null) continue;
      seen_null = 1;
      memset((char*)out + k * w, 0, (size_t)w);
      if (out_valid) out_valid[k >> 3] &= (uint8_t)~(1u << (k & 7));
      ++nulls; ++k;
      continue;
    }
    const uint64_t key = raw_key_at(bit_width, vals, off + i);
    if (rset_find(&s, key) >= 0) continue;
    rset_add(&s, key, i);
    memcpy((char*)out + k * w, (const char*)vals + (off + i) * w, (size_t)w);
    if (out_valid) out_valid[k >> 3] |= (uint8_t)(1u << (k & 7));
    ++k;
  }
  rset_free(&s);
  *out_len = k;
  if (out_nulls) *out_nulls = nulls;
  return REF_OK;
}
