// Package arrowgpu binds libarrowgpu (include/arrowgpu.h) into arrow-go: B200 implementations of the scalar compute
// kernels (arithmetic, comparison, boolean), filter / take and the arrow/math Sum reductions, registered in a child
// FunctionRegistry so compute.CallFunction / compute.Add / compute.Filter / compute.Take and the Substrait evaluator
// resolve to them with the Datum / ArraySpan API unchanged.
//
// Intended location: github.com/apache/arrow-go/v18/arrow/compute/gpu (inside the compute tree, so that
// arrow/compute/internal/kernels — option / state types — is importable).
//
// NOT BUILT IN THIS REPOSITORY: the build image has no Go toolchain.  The files are complete and were reviewed line by
// line against the reference (arrow/compute/exec/span.go:76-88, exec/utils.go:38-44, functions.go:233-310,
// registry.go:64-133, exec/kernel.go:617-727); every C entry point they call is exercised through the same C ABI by
// tests/ (ctypes, C consumer tests/c/abi_smoke.c) and by the C++ mirror of the executor in arrow_go_b200/host/.
// One `go test ./arrow/compute/...` with `ctx = compute.SetExecCtx(ctx, compute.ExecCtx{Registry: arrowgpu.NewRegistry()})`
// away from running the reference's own suites against the GPU registry.
package arrowgpu

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../arrow_go_b200/lib -larrowgpu -Wl,-rpath,${SRCDIR}/../../arrow_go_b200/lib
#include "arrowgpu.h"
#include <stdlib.h>
*/
import "C"

import (
	"fmt"
	"unsafe"

	"github.com/apache/arrow-go/v18/arrow"
	"github.com/apache/arrow-go/v18/arrow/compute/exec"
	"github.com/apache/arrow-go/v18/arrow/scalar"
)

// check maps an ag_status onto arrow-go's error sentinels (arrow/errors.go:21-28).  The message is the one the
// reference words ("overflow", "divide by zero", "%d out of bounds", ...).
func check(st C.ag_status) error {
	if st == C.AG_OK {
		return nil
	}
	var buf [512]C.char
	C.ag_last_error(&buf[0], 512)
	msg := C.GoString(&buf[0])
	switch st {
	case C.AG_ERR_INVALID:
		return fmt.Errorf("%w: %s", arrow.ErrInvalid, msg)
	case C.AG_ERR_INDEX:
		return fmt.Errorf("%w: %s", arrow.ErrIndex, msg)
	case C.AG_ERR_NOT_IMPLEMENTED:
		return fmt.Errorf("%w: %s", arrow.ErrNotImplemented, msg)
	case C.AG_ERR_TYPE:
		return fmt.Errorf("%w: %s", arrow.ErrType, msg)
	}
	return fmt.Errorf("%w: arrowgpu: %s", arrow.ErrInvalid, msg)
}

// byteWidth of a fixed-width span type.
func byteWidth(dt arrow.DataType) int64 {
	return int64(dt.(arrow.FixedWidthDataType).Bytes())
}

// valuesAt is exec.GetSpanValues without the typed reinterpretation (exec/utils.go:38-44): Buffers[1] advanced by
// Offset elements.  Every C entry point that takes a VALUES pointer expects it advanced like this; entry points that
// take (buffer, offset) pairs — bitmaps, and the filter / take / checked families — get element 0 plus the offset
// (see base0).  An empty span has no backing array to point into.
func valuesAt(a *exec.ArraySpan) unsafe.Pointer {
	if a.Len == 0 || len(a.Buffers[1].Buf) == 0 {
		return nil
	}
	return unsafe.Pointer(&a.Buffers[1].Buf[a.Offset*byteWidth(a.Type)])
}

// base0 is element 0 of a buffer (nil for an absent one).
func base0(b []byte) unsafe.Pointer {
	if len(b) == 0 {
		return nil
	}
	return unsafe.Pointer(&b[0])
}

func bitmap(b []byte) *C.uint8_t { return (*C.uint8_t)(base0(b)) }

// validityOf returns the span's validity bitmap only when it may hold nulls (Nulls != 0): a bitmap that is known to be
// all set takes the no-nulls path, exactly like ArraySpan.MayHaveNulls (exec/span.go:105-107).
func validityOf(a *exec.ArraySpan) *C.uint8_t {
	if !a.MayHaveNulls() {
		return nil
	}
	return bitmap(a.Buffers[0].Buf)
}

// scalarBytes is the raw little-endian value of a primitive scalar (scalar.PrimitiveScalar.Data, scalar.go:143-146),
// the `*(T*)scalar` the native loops read (base_arithmetic.cc:465-475).
func scalarBytes(s scalar.Scalar) (unsafe.Pointer, error) {
	p, ok := s.(scalar.PrimitiveScalar)
	if !ok {
		return nil, fmt.Errorf("%w: arrowgpu: scalar of type %s has no fixed-width representation", arrow.ErrNotImplemented, s.DataType())
	}
	d := p.Data()
	if len(d) == 0 {
		return nil, fmt.Errorf("%w: arrowgpu: empty scalar payload", arrow.ErrInvalid)
	}
	return unsafe.Pointer(&d[0]), nil
}

// shape of a binary batch: AG_SHAPE_AA / AS / SA.
func shapeOf(batch *exec.ExecSpan) C.int {
	switch {
	case batch.Values[0].IsArray() && batch.Values[1].IsArray():
		return C.AG_SHAPE_AA
	case batch.Values[0].IsArray():
		return C.AG_SHAPE_AS
	default:
		return C.AG_SHAPE_SA
	}
}

func boolInt(b bool) C.int {
	if b {
		return 1
	}
	return 0
}

// Version reports the library version string and the number of kernels this process has launched.
func Version() (string, uint64) { return C.GoString(C.ag_version()), uint64(C.ag_kernel_launch_count()) }
