package arrowgpu_test

// Smoke tests of the binding itself; the parity suite is the reference's own: run
//   go test ./arrow/compute/... -run 'TestAdd|TestSub|TestMultiply|NumericCompare|TestFilterNumeric|TestTakeNumeric'
// with every test context wrapped by compute.SetExecCtx(ctx, compute.ExecCtx{Registry: arrowgpu.NewRegistry(), ...}).

import (
	"context"
	"math"
	"strings"
	"testing"

	"github.com/apache/arrow-go/v18/arrow"
	"github.com/apache/arrow-go/v18/arrow/array"
	"github.com/apache/arrow-go/v18/arrow/compute"
	arrowgpu "github.com/apache/arrow-go/v18/arrow/compute/gpu"
	"github.com/apache/arrow-go/v18/arrow/memory"
)

func gpuCtx() context.Context {
	e := compute.DefaultExecCtx()
	e.Registry = arrowgpu.NewRegistry()
	return compute.SetExecCtx(context.Background(), e)
}

func TestAddOverflowAndValues(t *testing.T) { // arithmetic_test.go:325-359
	mem := memory.NewCheckedAllocator(arrowgpu.PinnedAllocator{})
	defer mem.AssertSize(t, 0)
	l, _, _ := array.FromJSON(mem, arrow.PrimitiveTypes.Int8, strings.NewReader(`[1, null, 127, -100]`))
	r, _, _ := array.FromJSON(mem, arrow.PrimitiveTypes.Int8, strings.NewReader(`[1, 5, 0, -1]`))
	defer l.Release()
	defer r.Release()
	ctx := gpuCtx()
	out, err := compute.Add(ctx, compute.ArithmeticOptions{}, compute.NewDatum(l), compute.NewDatum(r))
	if err != nil {
		t.Fatal(err)
	}
	defer out.Release()
	want, _, _ := array.FromJSON(mem, arrow.PrimitiveTypes.Int8, strings.NewReader(`[2, null, 127, -101]`))
	defer want.Release()
	if got := out.(*compute.ArrayDatum).MakeArray(); !array.Equal(got, want) {
		t.Fatalf("add: got %v want %v", got, want)
	}
	m, _, _ := array.FromJSON(mem, arrow.PrimitiveTypes.Int8, strings.NewReader(`[127]`))
	defer m.Release()
	if _, err := compute.Add(ctx, compute.ArithmeticOptions{}, compute.NewDatum(m), compute.NewDatum(m)); err == nil || !strings.Contains(err.Error(), "overflow") {
		t.Fatalf("expected overflow, got %v", err)
	}
}

func TestSumSiblings(t *testing.T) { // arrow/math/float64_test.go:30-48
	mem := arrowgpu.PinnedAllocator{}
	b := array.NewFloat64Builder(mem)
	defer b.Release()
	for i := 0; i < 10000; i++ {
		b.Append(float64(i))
	}
	a := b.NewFloat64Array()
	defer a.Release()
	if got := arrowgpu.SumFloat64(a); got != 49995000 {
		t.Fatalf("sum = %v", got)
	}
	if got := arrowgpu.SumFloat64ReferenceOrder(a); got != 49995000 || math.IsNaN(got) {
		t.Fatalf("reference-order sum = %v", got)
	}
}

func TestIsInUniqueSortIndices(t *testing.T) { // scalar_set_lookup_test.go:104-167, vector_hash_test.go:236-255, vector_sort_test.go:40-110
	mem := memory.NewCheckedAllocator(arrowgpu.PinnedAllocator{})
	defer mem.AssertSize(t, 0)
	ctx := gpuCtx()
	vals, _, _ := array.FromJSON(mem, arrow.PrimitiveTypes.Int32, strings.NewReader(`[2, 1, 2, 1, 2, 3, null]`))
	set, _, _ := array.FromJSON(mem, arrow.PrimitiveTypes.Int32, strings.NewReader(`[2, 3, null]`))
	defer vals.Release()
	defer set.Release()

	in, err := compute.IsIn(ctx, compute.SetOptions{ValueSet: compute.NewDatum(set)}, compute.NewDatum(vals))
	if err != nil {
		t.Fatal(err)
	}
	defer in.Release()
	wantIn, _, _ := array.FromJSON(mem, arrow.FixedWidthTypes.Boolean, strings.NewReader(`[true, false, true, false, true, true, true]`))
	defer wantIn.Release()
	if got := in.(*compute.ArrayDatum).MakeArray(); !array.Equal(got, wantIn) {
		t.Fatalf("is_in: got %v want %v", got, wantIn)
	}

	uq, err := compute.Unique(ctx, compute.NewDatum(vals))
	if err != nil {
		t.Fatal(err)
	}
	defer uq.Release()
	wantUq, _, _ := array.FromJSON(mem, arrow.PrimitiveTypes.Int32, strings.NewReader(`[2, 1, 3, null]`))
	defer wantUq.Release()
	if got := uq.(*compute.ArrayDatum).MakeArray(); !array.Equal(got, wantUq) {
		t.Fatalf("unique: got %v want %v", got, wantUq)
	}

	key := compute.DefaultSortKey()
	key.Order = compute.SortOrderDescending
	si, err := compute.CallFunction(ctx, "sort_indices", compute.SortOptions{key}, compute.NewDatum(vals))
	if err != nil {
		t.Fatal(err)
	}
	defer si.Release()
	wantSi, _, _ := array.FromJSON(mem, arrow.PrimitiveTypes.Uint64, strings.NewReader(`[5, 0, 2, 4, 1, 3, 6]`)) // stable, nulls at end
	defer wantSi.Release()
	if got := si.(*compute.ArrayDatum).MakeArray(); !array.Equal(got, wantSi) {
		t.Fatalf("sort_indices: got %v want %v", got, wantSi)
	}
}
