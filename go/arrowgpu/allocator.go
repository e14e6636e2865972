package arrowgpu

/*
#include "arrowgpu.h"
*/
import "C"

import (
	"unsafe"

	"github.com/apache/arrow-go/v18/arrow/memory"
)

// PinnedAllocator is a memory.Allocator (arrow/memory/allocator.go:23-27) over ag_host_alloc: page-locked, 64-byte
// aligned, zero-initialised C memory on the NUMA node of the calling thread's current GPU.  Arrays built with it are
// DMA'd straight from their buffers by the host-pointer entry points (no staging copy), can be exported as
// ARROW_DEVICE_CUDA_HOST and read in place by the *_dev entry points, and — being C memory — may be retained by C
// across calls, the same move as mallocator.Mallocator (arrow/memory/mallocator/mallocator.go:63-104).
type PinnedAllocator struct{}

var _ memory.Allocator = PinnedAllocator{}

func (PinnedAllocator) Allocate(size int) []byte {
	if size == 0 {
		return []byte{}
	}
	var p unsafe.Pointer
	if C.ag_host_alloc(&p, C.size_t(size)) != C.AG_OK {
		panic("arrowgpu: pinned allocation failed")
	}
	return unsafe.Slice((*byte)(p), size)
}

func (a PinnedAllocator) Reallocate(size int, b []byte) []byte {
	if len(b) == 0 {
		return a.Allocate(size)
	}
	if size == 0 {
		a.Free(b)
		return []byte{}
	}
	p := unsafe.Pointer(&b[0])
	if C.ag_host_realloc(&p, C.size_t(len(b)), C.size_t(size)) != C.AG_OK {
		panic("arrowgpu: pinned reallocation failed")
	}
	return unsafe.Slice((*byte)(p), size)
}

func (PinnedAllocator) Free(b []byte) {
	if len(b) == 0 {
		return
	}
	C.ag_host_free(unsafe.Pointer(&b[0]))
}
