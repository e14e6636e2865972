package arrowgpu

/*
#include "arrowgpu.h"
*/
import "C"

import (
	"runtime"
	"sync"
	"unsafe"

	"github.com/apache/arrow-go/v18/arrow/array"
)

// One Go process drives every GPU of the box: the reference's own fan-out is an errgroup of goroutines inside one
// process (arrow/compute/selection.go:127-150), so the binding does the same — one goroutine per device, each locked
// to an OS thread for the duration of its shard (the library tracks the current device per thread).

// InitAll initialises every visible device and enables peer access; returns the device count.
func InitAll() (int, error) {
	var n C.int
	if err := check(C.ag_init_all(&n)); err != nil {
		return 0, err
	}
	return int(n), nil
}

// OnDevice runs f on a goroutine pinned to an OS thread whose current device is `dev`.
func OnDevice(dev int, f func() error) error {
	errc := make(chan error, 1)
	go func() {
		runtime.LockOSThread()
		defer runtime.UnlockOSThread()
		if err := check(C.ag_set_device(C.int(dev))); err != nil {
			errc <- err
			return
		}
		errc <- f()
	}()
	return <-errc
}

// ShardRange is the row-range rule of the C ABI: ceil-balanced, cut at multiples of 64 rows.
func ShardRange(rows int64, shard, shards int) (start, stop int64) {
	var a, b C.int64_t
	if err := check(C.ag_shard_range(C.int64_t(rows), C.int(shard), C.int(shards), &a, &b)); err != nil {
		panic(err)
	}
	return int64(a), int64(b)
}

// SumInt64Sharded is BASELINE config 5 from Go: the column is cut by ShardRange, shard k is uploaded to device k
// (pinned source when the array was built with PinnedAllocator), each device runs the Sum kernel whose last block
// folds the partial results of all devices through HBM mailboxes over NVLink, and every device ends up with the
// global wrapping sum; device 0's copy is returned.
func SumInt64Sharded(a *array.Int64, devices int) (int64, error) {
	v := a.Int64Values()
	if len(v) == 0 {
		return 0, nil
	}
	devs := make([]C.int, devices)
	for i := range devs {
		devs[i] = C.int(i)
	}
	comms := make([]C.ag_comm_t, devices)
	if err := check(C.ag_comm_create_local(&comms[0], C.int(devices), &devs[0])); err != nil {
		return 0, err
	}
	results := make([]int64, devices)
	errs := make([]error, devices)
	var wg sync.WaitGroup
	for k := 0; k < devices; k++ {
		wg.Add(1)
		go func(k int) {
			defer wg.Done()
			errs[k] = OnDevice(k, func() error {
				lo, hi := ShardRange(int64(len(v)), k, devices)
				n := C.size_t(hi - lo)
				var dIn, dRes unsafe.Pointer
				if err := check(C.ag_dev_alloc(&dIn, n*8+8)); err != nil {
					return err
				}
				defer C.ag_dev_free(dIn)
				if err := check(C.ag_dev_alloc(&dRes, 8)); err != nil {
					return err
				}
				defer C.ag_dev_free(dRes)
				if n > 0 {
					if err := check(C.ag_upload(dIn, unsafe.Pointer(&v[lo]), n*8, nil)); err != nil {
						return err
					}
				}
				if err := check(C.ag_sum_i64_global_dev(comms[k], (*C.int64_t)(dIn), n, (*C.int64_t)(dRes), nil)); err != nil {
					return err
				}
				var r C.int64_t
				if err := check(C.ag_download(unsafe.Pointer(&r), dRes, 8, nil)); err != nil {
					return err
				}
				if err := check(C.ag_stream_sync(nil)); err != nil {
					return err
				}
				results[k] = int64(r)
				return nil
			})
		}(k)
	}
	wg.Wait()
	for k := range comms {
		C.ag_comm_destroy(comms[k])
	}
	for _, err := range errs {
		if err != nil {
			return 0, err
		}
	}
	return results[0], nil
}
