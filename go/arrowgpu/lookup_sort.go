package arrowgpu

/*
#include "arrowgpu.h"
*/
import "C"

import (
	"context"
	"fmt"

	"github.com/apache/arrow-go/v18/arrow"
	"github.com/apache/arrow-go/v18/arrow/array"
	"github.com/apache/arrow-go/v18/arrow/compute"
	"github.com/apache/arrow-go/v18/arrow/compute/exec"
	"github.com/apache/arrow-go/v18/arrow/compute/internal/kernels"
	"github.com/apache/arrow-go/v18/arrow/memory"
)

// ---- is_in (scalar_set_lookup.go:70-171, kernels/scalar_set_lookup.go:112-413) ---------------------------------
//
// The reference's kernel state is a memo table of the value set; here it is the value set itself as ONE array of the
// argument's type (cast with SafeCastOptions and concatenated exactly where initSetLookup does it), and the table is
// built on the device inside ag_is_in.  Values are compared by their raw bytes on both sides, like the memo tables.

type setState struct {
	set      arrow.Array
	behavior C.int // NullMatchingBehavior: MATCH 0, SKIP 1, EMIT_NULL 2, INCONCLUSIVE 3 (= AG_NULL_*)
}

func initIsIn(ctx *exec.KernelCtx, args exec.KernelInitArgs) (exec.KernelState, error) {
	if args.Options == nil {
		return nil, fmt.Errorf("%w: calling a set lookup function without SetOptions", compute.ErrInvalid)
	}
	opts, ok := args.Options.(*compute.SetOptions)
	if !ok {
		return nil, fmt.Errorf("%w: expected SetOptions, got %T", compute.ErrInvalid, args.Options)
	}
	valueset, ok := opts.ValueSet.(compute.ArrayLikeDatum)
	if !ok {
		return nil, fmt.Errorf("%w: expected array-like datum, got %T", compute.ErrInvalid, opts.ValueSet)
	}
	argType := args.Inputs[0]
	var owned compute.Datum
	if !arrow.TypeEqual(valueset.Type(), argType) { // scalar_set_lookup.go:94-112: the set is cast to the argument's type
		result, err := compute.CastDatum(ctx.Ctx, valueset, compute.SafeCastOptions(argType))
		if err != nil {
			return nil, fmt.Errorf("%w: array type doesn't match type of values set: %s vs %s", compute.ErrInvalid, argType, valueset.Type())
		}
		owned = result
		valueset = result.(compute.ArrayLikeDatum)
	}
	if owned != nil {
		defer owned.Release()
	}
	var set arrow.Array
	switch valueset.Kind() {
	case compute.KindArray:
		set = valueset.(*compute.ArrayDatum).MakeArray()
	case compute.KindChunked:
		var err error
		if set, err = array.Concatenate(valueset.(*compute.ChunkedDatum).Value.Chunks(), exec.GetAllocator(ctx.Ctx)); err != nil {
			return nil, err
		}
	default:
		return nil, fmt.Errorf("%w: expected array or chunked array, got %s", compute.ErrInvalid, opts.ValueSet.Kind())
	}
	return &setState{set: set, behavior: C.int(opts.NullBehavior)}, nil
}

// isInExec writes the data and validity bitmaps of one batch (isInKernelExec :373-413).  The kernel is registered with
// CanWriteIntoSlices = false, so `out` is a fresh allocation at bit offset 0 — what the C entry point expects.
func isInExec(ctx *exec.KernelCtx, batch *exec.ExecSpan, out *exec.ExecResult) error {
	st := ctx.State.(*setState)
	in := &batch.Values[0].Array
	if out.Offset != 0 {
		return fmt.Errorf("%w: arrowgpu: is_in into a sliced output", arrow.ErrNotImplemented)
	}
	set := st.set.Data()
	var setValid *C.uint8_t
	if set.NullN() != 0 && set.Buffers()[0] != nil {
		setValid = bitmap(set.Buffers()[0].Bytes())
	}
	var setVals []byte
	if set.Len() > 0 {
		setVals = set.Buffers()[1].Bytes()
	}
	bw := in.Type.(arrow.FixedWidthDataType).BitWidth()
	var nulls C.int64_t
	err := check(C.ag_is_in(C.int(bw), base0(in.Buffers[1].Buf), validityOf(in), C.int64_t(in.Offset), C.int64_t(in.Len),
		base0(setVals), setValid, C.int64_t(set.Offset()), C.int64_t(set.Len()), st.behavior,
		bitmap(out.Buffers[1].Buf), bitmap(out.Buffers[0].Buf), &nulls))
	if err != nil {
		return err
	}
	out.Nulls = int64(nulls)
	return nil
}

func isInFunction() *compute.ScalarFunction {
	fn := compute.NewScalarFunction("is_in", compute.Unary(), compute.EmptyFuncDoc)
	for _, ty := range numericTypes {
		kn := exec.NewScalarKernel([]exec.InputType{exec.NewExactInput(ty)}, exec.NewOutputType(arrow.FixedWidthTypes.Boolean), isInExec, initIsIn)
		kn.MemAlloc = exec.MemPrealloc                 // scalar_set_lookup.go:190-191
		kn.NullHandling = exec.NullComputedPrealloc
		kn.CanWriteIntoSlices = false
		kn.CleanupFn = func(state exec.KernelState) error {
			if s, ok := state.(*setState); ok && s.set != nil {
				s.set.Release()
				s.set = nil
			}
			return nil
		}
		if err := fn.AddKernel(kn); err != nil {
			panic(err)
		}
	}
	return fn
}

// ---- unique (vector_hash.go:105-113, kernels/vector_hash.go) -------------------------------------------------------
//
// The reference hashes chunk by chunk into one memo table and materialises it in Finalize; ag_unique needs the whole
// column at once (the first row of every key is a global minimum), so the kernel is not chunk-wise: a chunked input is
// concatenated first.  Output: the distinct values in order of first appearance, one null where it first appears.

func uniqueExec(ctx *exec.KernelCtx, batch *exec.ExecSpan, out *exec.ExecResult) error {
	in := &batch.Values[0].Array
	bw := in.Type.(arrow.FixedWidthDataType).BitWidth()
	preallocate(ctx, in.Len, bw, in.MayHaveNulls(), out) // n slots are always enough; Len is trimmed below
	var outLen, outNulls C.int64_t
	err := check(C.ag_unique(C.int(bw), base0(in.Buffers[1].Buf), validityOf(in), C.int64_t(in.Offset), C.int64_t(in.Len),
		base0(out.Buffers[1].Buf), bitmap(out.Buffers[0].Buf), &outLen, &outNulls))
	if err != nil {
		return err
	}
	out.Len, out.Nulls = int64(outLen), int64(outNulls)
	return nil
}

func uniqueChunked(ctx *exec.KernelCtx, cols []*arrow.Chunked, out *exec.ExecResult) ([]*exec.ExecResult, error) {
	arr, err := array.Concatenate(cols[0].Chunks(), exec.GetAllocator(ctx.Ctx))
	if err != nil {
		return nil, err
	}
	defer arr.Release()
	span := exec.ExecSpan{Len: int64(arr.Len()), Values: make([]exec.ExecValue, 1)}
	span.Values[0].Array.SetMembers(arr.Data())
	if err := uniqueExec(ctx, &span, out); err != nil {
		return nil, err
	}
	return []*exec.ExecResult{out}, nil
}

func uniqueFunction() *compute.VectorFunction {
	fn := compute.NewVectorFunction("unique", compute.Unary(), compute.EmptyFuncDoc)
	for _, ty := range numericTypes {
		k := exec.NewVectorKernel([]exec.InputType{exec.NewExactInput(ty)}, kernels.OutputFirstType, uniqueExec, nil)
		k.CanExecuteChunkWise = false
		k.ExecChunked = uniqueChunked
		k.NullHandling = exec.NullComputedNoPrealloc
		k.MemAlloc = exec.MemNoPrealloc
		if err := fn.AddKernel(k); err != nil {
			panic(err)
		}
	}
	return fn
}

// ---- sort_indices (vector_sort.go:42-204, kernels.SortIndices vector_sort.go:385-481) -------------------------------
//
// A MetaFunction in the reference.  The GPU entry sorts ONE contiguous fixed-width column (stable, NaNs after the finite
// values, nulls per NullPlacement); chunked inputs, record batches, tables and multi-key sorts go to the parent's
// implementation, like every other shape this package does not cover.

func sortIndicesFunction(parent compute.Function) *compute.MetaFunction {
	return compute.NewMetaFunction("sort_indices", compute.Unary(), compute.EmptyFuncDoc,
		func(ctx context.Context, opts compute.FunctionOptions, args ...compute.Datum) (compute.Datum, error) {
			keys, isKeys := opts.(compute.SortOptions)
			if in, ok := args[0].(*compute.ArrayDatum); ok && isKeys && len(keys) >= 1 && isNumeric(in.Type()) {
				return sortOneColumn(ctx, in.Value, keys[0]) // a bare array uses the first key only (vector_sort.go:138-142)
			}
			if parent == nil {
				return nil, fmt.Errorf("%w: unsupported type for sort_indices operation: %s", arrow.ErrNotImplemented, args[0])
			}
			return parent.Execute(ctx, opts, args...)
		})
}

func isNumeric(dt arrow.DataType) bool {
	for _, ty := range numericTypes {
		if arrow.TypeEqual(ty, dt) {
			return true
		}
	}
	return false
}

func sortOneColumn(ctx context.Context, col arrow.ArrayData, key compute.SortKey) (compute.Datum, error) {
	n := col.Len()
	buf := memory.NewResizableBuffer(exec.GetAllocator(ctx))
	buf.Resize(n * 8)
	var vals []byte
	if n > 0 {
		vals = col.Buffers()[1].Bytes()
	}
	var valid *C.uint8_t
	if col.NullN() != 0 && col.Buffers()[0] != nil {
		valid = bitmap(col.Buffers()[0].Bytes())
	}
	var nulls, nans C.int64_t
	// arrow.Type ids are the AG_TYPE_* ids; SortOrder / NullPlacement are 0 / 1 on both sides (kernels/vector_sort.go:34-46)
	err := check(C.ag_sort_indices(C.int(col.DataType().ID()), base0(vals), valid, C.int64_t(col.Offset()), C.int64_t(n),
		C.int(key.Order), C.int(key.NullPlacement), (*C.uint64_t)(base0(buf.Bytes())), &nulls, &nans))
	if err != nil {
		buf.Release()
		return nil, err
	}
	out := array.NewData(arrow.PrimitiveTypes.Uint64, n, []*memory.Buffer{nil, buf}, nil, 0, 0)
	buf.Release() // NewData retained it
	return &compute.ArrayDatum{Value: out}, nil
}

// registerLookupSort adds is_in, unique and sort_indices to a registry built by NewRegistry.
func registerLookupSort(reg, parent compute.FunctionRegistry) {
	addScalar(reg, parent, isInFunction())
	addVector(reg, parent, uniqueFunction())
	p, _ := parent.GetFunction("sort_indices")
	reg.AddFunction(sortIndicesFunction(p), true)
}
