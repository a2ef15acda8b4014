// +build gpu

// Package chunk — bridge between chunk.Column and the tq_column of libtinysql_b200 (include/tinysql_b200.h).
// A chunk.Column (column.go:28-34) is already laid out the way the C-ABI wants it: data (8-byte slots, 4-byte FLOAT
// slots, or var-len bytes), offsets (var-len only) and a NOT-NULL bitmap with bit (i&7) of byte (i>>3).
package chunk

/*
#cgo CFLAGS: -I${SRCDIR}/../../../../include
#cgo LDFLAGS: -ltinysql_b200
#include <stdlib.h>
#include <string.h>
#include "tinysql_b200.h"
*/
import "C"

import (
	"reflect"
	"unsafe"
)

// CColumn is a tq_column view of a Column.
type CColumn = C.tq_column

// CViewSet is the argument block of one C-ABI call: an array of tq_column structs plus the buffers they point to.
//
// cgo pointer rules (cmd/cgo "Passing pointers"): Go memory handed to C must not itself contain Go pointers, and C must
// not keep a Go pointer after the call.  A []tq_column in Go memory whose fields point at Go slices breaks the first rule
// (cgocheck panics with "cgo argument has Go pointer to unpinned Go pointer").  So the tq_column array lives in C memory
// (C.calloc), and every buffer it points to is either
//   (a) C memory already — columns allocated by NewPinnedColumn (tq_pinned_alloc): zero copy, and because nothing the
//       library could retain is a Go pointer, such columns may also be used with TQ_JOIN_STABLE_INPUT; or
//   (b) pinned for the duration of the call with runtime.Pinner (Go >= 1.21): zero copy, released by Release(); or
//   (c) copied into the set's C scratch area (older toolchains): one memcpy per column, the same cost as the library's
//       own staging copy of a <=1024-row chunk.
// A set is owned by one executor / builtin instance and reused across calls (no allocation in steady state).
type CViewSet struct {
	cols    *C.tq_column // C array of n entries
	n       int
	scratch []cBuf   // per column: data, offsets, bitmap staging (strategy c)
	pinner  pinner   // strategy (b); a no-op type on toolchains without runtime.Pinner
}

type cBuf struct {
	p   unsafe.Pointer
	cap int
}

func (b *cBuf) ensure(n int) unsafe.Pointer {
	if n > b.cap {
		if b.p != nil {
			C.free(b.p)
		}
		b.p, b.cap = C.malloc(C.size_t(n+n/2+64)), n+n/2+64
	}
	return b.p
}

// NewCViewSet allocates the argument block for n columns.
func NewCViewSet(n int) *CViewSet {
	return &CViewSet{cols: (*C.tq_column)(C.calloc(C.size_t(n), C.size_t(unsafe.Sizeof(C.tq_column{})))), n: n, scratch: make([]cBuf, 3*n)}
}

// Free releases the C memory of the set.
func (s *CViewSet) Free() {
	C.free(unsafe.Pointer(s.cols))
	for i := range s.scratch {
		if s.scratch[i].p != nil {
			C.free(s.scratch[i].p)
		}
	}
	s.cols = nil
}

// Ptr is what a tq_* call takes.
func (s *CViewSet) Ptr() *C.tq_column { return s.cols }

// At returns entry i of the C array (for entry points that take single tq_column pointers, the tq_vec_* family).
func (s *CViewSet) At(i int) *C.tq_column { return s.at(i) }

func (s *CViewSet) at(i int) *C.tq_column {
	return (*C.tq_column)(unsafe.Pointer(uintptr(unsafe.Pointer(s.cols)) + uintptr(i)*unsafe.Sizeof(C.tq_column{})))
}

// stage makes buffer `b` reachable from C for the coming call and returns the address to store in the tq_column.
func (s *CViewSet) stage(slot int, b []byte, cOwned bool) unsafe.Pointer {
	if len(b) == 0 {
		return nil
	}
	if cOwned { // (a)
		return unsafe.Pointer(&b[0])
	}
	if s.pinner.available() { // (b)
		s.pinner.Pin(&b[0])
		return unsafe.Pointer(&b[0])
	}
	dst := s.scratch[slot].ensure(len(b)) // (c)
	C.memcpy(dst, unsafe.Pointer(&b[0]), C.size_t(len(b)))
	return dst
}

// Fill points entry i at column c (input columns: the library only reads them).
func (s *CViewSet) Fill(i int, c *Column) {
	v := s.at(i)
	v.length = C.int64_t(c.length)
	v.null_bitmap = (*C.uint8_t)(s.stage(3*i, c.nullBitmap, false))
	v.offsets = nil
	if !c.isFixed() { // length+1 entries, offsets[0] == 0
		v.offsets = (*C.int64_t)(s.stage(3*i+1, int64Bytes(c.offsets), false))
	}
	v.data = (*C.uint8_t)(s.stage(3*i+2, c.data, c.cOwned))
}

// FillChunk fills one entry per column of the chunk.
func (s *CViewSet) FillChunk(chk *Chunk) {
	for i, col := range chk.columns {
		s.Fill(i, col)
	}
}

// Release ends the call: unpins what strategy (b) pinned.
func (s *CViewSet) Release() { s.pinner.Unpin() }

func int64Bytes(v []int64) []byte {
	if len(v) == 0 {
		return nil
	}
	return (*[1 << 40]byte)(unsafe.Pointer(&v[0]))[: len(v)*8 : len(v)*8]
}

// Result columns are written by the library, so strategy (c) needs a copy back: FillResult points entry i at C scratch
// sized for the coming call (or at the pinned / C-owned Go buffer), CopyBack moves the produced rows into the Column.
func (s *CViewSet) FillResult(i int, c *Column) {
	v := s.at(i)
	v.length = C.int64_t(c.length)
	direct := c.cOwned || s.pinner.available()
	pick := func(slot int, b []byte) unsafe.Pointer {
		if len(b) == 0 {
			return nil
		}
		if direct {
			return s.stage(slot, b, c.cOwned)
		}
		return s.scratch[slot].ensure(len(b))
	}
	v.null_bitmap = (*C.uint8_t)(pick(3*i, c.nullBitmap))
	v.offsets = nil
	if !c.isFixed() {
		v.offsets = (*C.int64_t)(pick(3*i+1, int64Bytes(c.offsets)))
	}
	v.data = (*C.uint8_t)(pick(3*i+2, c.data))
}

// CopyBack copies the first n produced rows of result entry i from C scratch into the Column (strategy c only).
func (s *CViewSet) CopyBack(i int, c *Column, n int) {
	if c.cOwned || s.pinner.available() || n == 0 {
		return
	}
	v := s.at(i)
	copyOut := func(dst []byte, src unsafe.Pointer, bytes int) {
		if bytes > 0 {
			C.memcpy(unsafe.Pointer(&dst[0]), src, C.size_t(bytes))
		}
	}
	copyOut(c.nullBitmap, unsafe.Pointer(v.null_bitmap), (n+7)>>3)
	if c.isFixed() {
		copyOut(c.data, unsafe.Pointer(v.data), n*len(c.elemBuf))
	} else {
		copyOut(int64Bytes(c.offsets), unsafe.Pointer(v.offsets), (n+1)*8)
		copyOut(c.data, unsafe.Pointer(v.data), int(c.offsets[n]))
	}
}

// PrepareFixedResult makes room for n rows of an 8-byte (elemLen 8) or FLOAT (elemLen 4) result column the way
// ResizeInt64(n, false) does (column.go:220-260,331) and returns the writable view; the callee fills data and bitmap.
func (c *Column) PrepareFixedResult(n, elemLen int) {
	c.resize(n, elemLen, false)
}

// PrepareVarLenResult sizes a var-len result column for n rows and dataBytes bytes of cells (the sizes come from
// tq_join_next_bytes) and returns the writable view.
func (c *Column) PrepareVarLenResult(n int, dataBytes int64) {
	c.reserve(n, 8)
	if int64(cap(c.data)) < dataBytes {
		c.data = make([]byte, dataBytes)
	}
	(*reflect.SliceHeader)(unsafe.Pointer(&c.data)).Len = int(dataBytes)
	if cap(c.offsets) < n+1 {
		c.offsets = make([]int64, n+1)
	}
	c.offsets = c.offsets[:n+1]
	sizeNulls := (n + 7) >> 3
	if cap(c.nullBitmap) < sizeNulls {
		c.nullBitmap = make([]byte, sizeNulls)
	}
	c.nullBitmap = c.nullBitmap[:sizeNulls]
	c.length = n
}

// SetResultRows trims a result column to the n rows the library actually produced.
func (c *Column) SetResultRows(n int) {
	c.length = n
	if c.isFixed() {
		c.data = c.data[:n*len(c.elemBuf)]
	} else {
		c.offsets = c.offsets[:n+1]
		c.data = c.data[:c.offsets[n]]
	}
	c.nullBitmap = c.nullBitmap[:(n+7)>>3]
}

// NewPinnedColumn allocates the data buffer of a fixed-width column from page-locked C memory (tq_pinned_alloc), so
// host<->device copies of the column are plain DMA and — when the column is not recycled while an operator reads it —
// the operator may be created with TQ_JOIN_STABLE_INPUT (uploads then overlap result downloads).  Release with
// FreePinned: the Go GC does not own this memory.
func NewPinnedColumn(elemLen, capRows int) (*Column, error) {
	var p unsafe.Pointer
	if st := C.tq_pinned_alloc(C.size_t(elemLen*capRows), &p); st != C.TQ_OK {
		return nil, StatusError(int32(st))
	}
	c := &Column{elemBuf: make([]byte, elemLen), nullBitmap: make([]byte, 0, (capRows+7)>>3), cOwned: true} // cOwned: a field this patch adds to Column
	hdr := (*reflect.SliceHeader)(unsafe.Pointer(&c.data))
	hdr.Data, hdr.Len, hdr.Cap = uintptr(p), 0, elemLen*capRows
	return c, nil
}

// FreePinned returns a NewPinnedColumn buffer.
func (c *Column) FreePinned() {
	if cap(c.data) > 0 {
		C.tq_pinned_free(unsafe.Pointer(&c.data[:1][0]))
		c.data = nil
	}
}

// StatusError turns a non-zero library status into an error carrying tq_last_error's text; the expression and executor
// packages map the reference-visible kinds (overflow, unsupported type) onto their own error values.
type StatusError int32

func (s StatusError) Error() string {
	buf := make([]byte, 512)
	C.tq_last_error((*C.char)(unsafe.Pointer(&buf[0])), C.int32_t(len(buf)))
	n := 0
	for n < len(buf) && buf[n] != 0 {
		n++
	}
	return string(buf[:n])
}
