package arrowgpu

/*
#include "arrowgpu.h"
*/
import "C"

import (
	"fmt"
	"unsafe"

	"github.com/apache/arrow-go/v18/arrow"
	"github.com/apache/arrow-go/v18/arrow/compute/exec"
	"github.com/apache/arrow-go/v18/arrow/compute/internal/kernels"
)

// ---- arithmetic, every slot computed: ScalarBinary (kernels/helpers.go:193-236) over the native loop, exactly the
// shape of base_arithmetic_avx2_amd64.go:35-39.  The executor has already intersected the validity bitmaps
// (NullIntersection, exec/kernel.go:457-476); out arrives preallocated, possibly as a slice of one contiguous buffer
// (out.Offset != 0), which valuesAt honours.
func arithExec(op C.int8_t) exec.ArrayKernelExec {
	return func(_ *exec.KernelCtx, batch *exec.ExecSpan, out *exec.ExecResult) error {
		typ := C.int(out.Type.ID())
		n := C.int64_t(batch.Len)
		l, r := &batch.Values[0], &batch.Values[1]
		switch shapeOf(batch) {
		case C.AG_SHAPE_AA:
			return check(C.ag_arith_binary(typ, op, valuesAt(&l.Array), valuesAt(&r.Array), valuesAt(out), n))
		case C.AG_SHAPE_AS:
			s, err := scalarBytes(r.Scalar)
			if err != nil {
				return err
			}
			return check(C.ag_arith_arr_scalar(typ, op, valuesAt(&l.Array), s, valuesAt(out), n))
		default:
			s, err := scalarBytes(l.Scalar)
			if err != nil {
				return err
			}
			return check(C.ag_arith_scalar_arr(typ, op, s, valuesAt(&r.Array), valuesAt(out), n))
		}
	}
}

// ---- checked integer add / sub / mul / div: ScalarBinaryNotNull (helpers.go:284-380) with the overflow predicate of
// base_arithmetic.go:249-294.  ag_arith_checked takes VALUE pointers already advanced by the span offset and
// (bitmap, bit offset) pairs for the validity; a scalar operand passes its value pointer with a NULL bitmap (a null
// scalar never reaches a kernel: the executor short-circuits all-null arguments, executor.go:237-349).
func checkedExec(op C.int8_t) exec.ArrayKernelExec {
	return func(_ *exec.KernelCtx, batch *exec.ExecSpan, out *exec.ExecResult) error {
		typ := C.int(out.Type.ID())
		shape := shapeOf(batch)
		var lp, rp unsafe.Pointer
		var lv, rv *C.uint8_t
		var lo, ro C.int64_t
		var err error
		if batch.Values[0].IsArray() {
			a := &batch.Values[0].Array
			lp, lv, lo = valuesAt(a), validityOf(a), C.int64_t(a.Offset)
		} else if lp, err = scalarBytes(batch.Values[0].Scalar); err != nil {
			return err
		}
		if batch.Values[1].IsArray() {
			a := &batch.Values[1].Array
			rp, rv, ro = valuesAt(a), validityOf(a), C.int64_t(a.Offset)
		} else if rp, err = scalarBytes(batch.Values[1].Scalar); err != nil {
			return err
		}
		var bad C.int64_t
		return check(C.ag_arith_checked(typ, op, shape, lp, lv, lo, rp, rv, ro, valuesAt(out), C.int64_t(batch.Len), &bad))
	}
}

// ---- unary abs / negate / sign (base_arithmetic.cc:412-438)
func unaryExec(op C.int8_t) exec.ArrayKernelExec {
	return func(_ *exec.KernelCtx, batch *exec.ExecSpan, out *exec.ExecResult) error {
		in := &batch.Values[0].Array
		if out.Type.ID() == in.Type.ID() {
			return check(C.ag_arith_unary_same(C.int(in.Type.ID()), op, valuesAt(in), valuesAt(out), C.int64_t(batch.Len)))
		}
		return check(C.ag_arith_unary_diff(C.int(in.Type.ID()), C.int(out.Type.ID()), op, valuesAt(in), valuesAt(out), C.int64_t(batch.Len)))
	}
}

// ---- comparisons: compareKernel (scalar_comparisons.go:199-218).  The output pointer is the byte that holds the first
// output bit and the prefix is out.Offset % 8; bits outside [offset, offset+len) are preserved.  less / less_equal are
// greater / greater_equal with the operands flipped (scalar_compare.go:73-99) — the C side does the flip for
// AG_CMP_LT / AG_CMP_LE.
func compareExec(cmp C.int) exec.ArrayKernelExec {
	return func(_ *exec.KernelCtx, batch *exec.ExecSpan, out *exec.ExecResult) error {
		var typ C.int
		var lp, rp unsafe.Pointer
		var err error
		if batch.Values[0].IsArray() {
			typ, lp = C.int(batch.Values[0].Array.Type.ID()), valuesAt(&batch.Values[0].Array)
		} else if lp, err = scalarBytes(batch.Values[0].Scalar); err != nil {
			return err
		}
		if batch.Values[1].IsArray() {
			typ, rp = C.int(batch.Values[1].Array.Type.ID()), valuesAt(&batch.Values[1].Array)
		} else if rp, err = scalarBytes(batch.Values[1].Scalar); err != nil {
			return err
		}
		if batch.Len == 0 {
			return nil
		}
		bits := (*C.uint8_t)(unsafe.Pointer(&out.Buffers[1].Buf[out.Offset/8]))
		return check(C.ag_compare(typ, cmp, shapeOf(batch), lp, rp, bits, C.int64_t(batch.Len), C.int(out.Offset%8)))
	}
}

// ---- boolean kernels on data bitmaps: and / or / xor / and_not (validity by intersection, done by the executor) —
// scalar_boolean.go:67-140; bitmaps are (buffer, bit offset) pairs.
func bitmapOpExec(op C.int) exec.ArrayKernelExec {
	return func(_ *exec.KernelCtx, batch *exec.ExecSpan, out *exec.ExecResult) error {
		if !batch.Values[0].IsArray() || !batch.Values[1].IsArray() {
			return fmt.Errorf("%w: arrowgpu: boolean kernels with a scalar operand are left to the built-in function", arrow.ErrNotImplemented)
		}
		l, r := &batch.Values[0].Array, &batch.Values[1].Array
		return check(C.ag_bitmap_op(op, bitmap(l.Buffers[1].Buf), C.int64_t(l.Offset), bitmap(r.Buffers[1].Buf), C.int64_t(r.Offset),
			bitmap(out.Buffers[1].Buf), C.int64_t(out.Offset), C.int64_t(batch.Len)))
	}
}

// Kleene and / or / and_not (scalar_boolean.go:142-347): NullComputedPrealloc — the kernel writes both the data and
// the validity bitmap of out.  A missing input validity is passed as NULL (all valid).
func kleeneExec(op C.int) exec.ArrayKernelExec {
	return func(_ *exec.KernelCtx, batch *exec.ExecSpan, out *exec.ExecResult) error {
		if !batch.Values[0].IsArray() || !batch.Values[1].IsArray() {
			return fmt.Errorf("%w: arrowgpu: Kleene kernels with a scalar operand are left to the built-in function", arrow.ErrNotImplemented)
		}
		l, r := &batch.Values[0].Array, &batch.Values[1].Array
		st := C.ag_kleene(op, bitmap(l.Buffers[1].Buf), validityOf(l), C.int64_t(l.Offset), bitmap(r.Buffers[1].Buf), validityOf(r), C.int64_t(r.Offset),
			bitmap(out.Buffers[1].Buf), bitmap(out.Buffers[0].Buf), C.int64_t(out.Offset), C.int64_t(batch.Len))
		out.Nulls = -1 // array.UnknownNullCount: recounted by the executor (executor.go:706-708)
		return check(st)
	}
}

// ---- PrimitiveFilter (vector_selection.go:449-520) = getFilterOutputSize + preallocateData + compaction.
func preallocate(ctx *exec.KernelCtx, length int64, bitWidth int, validity bool, out *exec.ExecResult) { // :83-93
	out.Len = length
	if validity {
		out.Buffers[0].WrapBuffer(ctx.AllocateBitmap(length))
	}
	if bitWidth == 1 {
		out.Buffers[1].WrapBuffer(ctx.AllocateBitmap(length))
	} else {
		out.Buffers[1].WrapBuffer(ctx.Allocate(int(length) * (bitWidth / 8)))
	}
}

func filterExec(ctx *exec.KernelCtx, batch *exec.ExecSpan, out *exec.ExecResult) error {
	values, filter := &batch.Values[0].Array, &batch.Values[1].Array
	if values.Len != filter.Len {
		return fmt.Errorf("%w: values and filter must have identical lengths", arrow.ErrInvalid)
	}
	sel := C.int(ctx.State.(kernels.FilterState).NullSelection)
	var n C.int64_t
	if err := check(C.ag_filter_output_size(bitmap(filter.Buffers[1].Buf), validityOf(filter), C.int64_t(filter.Offset),
		C.int64_t(filter.Len), sel, &n)); err != nil {
		return err
	}
	bw := values.Type.(arrow.FixedWidthDataType).BitWidth()
	withValidity := values.MayHaveNulls() || filter.MayHaveNulls() // :473
	preallocate(ctx, int64(n), bw, withValidity, out)
	var outLen, outNulls C.int64_t
	// (buffer element 0, element offset) pairs: the C side advances by offset * width itself
	err := check(C.ag_filter_primitive(C.int(bw), base0(values.Buffers[1].Buf), validityOf(values), C.int64_t(values.Offset),
		bitmap(filter.Buffers[1].Buf), validityOf(filter), C.int64_t(filter.Offset), C.int64_t(values.Len), sel,
		base0(out.Buffers[1].Buf), bitmap(out.Buffers[0].Buf), &outLen, &outNulls))
	if err != nil {
		return err
	}
	out.Nulls = int64(outNulls)
	return nil
}

// ---- PrimitiveTake (vector_selection.go:1162-1192) = checkIndexBounds + preallocateData + gather.
func takeExec(ctx *exec.KernelCtx, batch *exec.ExecSpan, out *exec.ExecResult) error {
	values, indices := &batch.Values[0].Array, &batch.Values[1].Array
	bw := values.Type.(arrow.FixedWidthDataType).BitWidth()
	preallocate(ctx, indices.Len, bw, values.MayHaveNulls() || indices.MayHaveNulls(), out) // :1175
	idxBits := C.int(indices.Type.(arrow.FixedWidthDataType).BitWidth())
	var nulls, badPos, badIdx C.int64_t
	err := check(C.ag_take_primitive(C.int(bw), base0(values.Buffers[1].Buf), validityOf(values), C.int64_t(values.Offset), C.int64_t(values.Len),
		idxBits, boolInt(arrow.IsSignedInteger(indices.Type.ID())), valuesAt(indices), validityOf(indices), C.int64_t(indices.Offset),
		C.int64_t(indices.Len), boolInt(ctx.State.(kernels.TakeState).BoundsCheck),
		base0(out.Buffers[1].Buf), bitmap(out.Buffers[0].Buf), &nulls, &badPos, &badIdx))
	if err != nil {
		return err // ErrIndex: "%d out of bounds" (helpers.go:951)
	}
	out.Nulls = int64(nulls)
	return nil
}

// ---- numeric cast with the safe-cast checks (numeric_cast.go:37-71, helpers.go:496-653)
func castExec(ctx *exec.KernelCtx, batch *exec.ExecSpan, out *exec.ExecResult) error {
	opts := ctx.State.(kernels.CastState)
	in := &batch.Values[0].Array
	var bad C.int64_t
	return check(C.ag_cast_numeric_checked(C.int(in.Type.ID()), C.int(out.Type.ID()), valuesAt(in), validityOf(in), C.int64_t(in.Offset),
		valuesAt(out), C.int64_t(in.Len), boolInt(opts.AllowIntOverflow), boolInt(opts.AllowFloatTruncate), &bad))
}
