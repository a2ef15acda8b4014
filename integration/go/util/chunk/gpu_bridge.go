// +build gpu

// Package chunk — bridge between chunk.Column and the tq_column of libtinysql_b200 (include/tinysql_b200.h).
// A chunk.Column (column.go:28-34) is already laid out the way the C-ABI wants it: data (8-byte slots, 4-byte FLOAT
// slots, or var-len bytes), offsets (var-len only) and a NOT-NULL bitmap with bit (i&7) of byte (i>>3).
package chunk

/*
#cgo CFLAGS: -I${SRCDIR}/../../../../include
#cgo LDFLAGS: -ltinysql_b200
#include "tinysql_b200.h"
*/
import "C"

import (
	"reflect"
	"unsafe"
)

// CColumn is a tq_column view of a Column. It borrows the Go buffers: valid only for the duration of ONE cgo call
// (the library copies what it keeps before returning, unless the handle was created with TQ_JOIN_STABLE_INPUT and the
// column lives in pinned C memory, see NewPinnedColumn).
type CColumn = C.tq_column

// CView fills v with c's buffers. No copy.
func (c *Column) CView(v *CColumn) {
	v.length = C.int64_t(c.length)
	v.null_bitmap = nil
	v.offsets = nil
	v.data = nil
	if len(c.nullBitmap) > 0 {
		v.null_bitmap = (*C.uint8_t)(unsafe.Pointer(&c.nullBitmap[0]))
	}
	if !c.isFixed() {
		v.offsets = (*C.int64_t)(unsafe.Pointer(&c.offsets[0])) // length+1 entries, offsets[0] == 0
	}
	if len(c.data) > 0 {
		v.data = (*C.uint8_t)(unsafe.Pointer(&c.data[0]))
	}
}

// CViews fills one tq_column per column of the chunk.
func (chk *Chunk) CViews(vs []CColumn) {
	for i, col := range chk.columns {
		col.CView(&vs[i])
	}
}

// PrepareFixedResult makes room for n rows of an 8-byte (elemLen 8) or FLOAT (elemLen 4) result column the way
// ResizeInt64(n, false) does (column.go:220-260,331) and returns the writable view; the callee fills data and bitmap.
func (c *Column) PrepareFixedResult(n, elemLen int, v *CColumn) {
	c.resize(n, elemLen, false)
	c.CView(v)
}

// PrepareVarLenResult sizes a var-len result column for n rows and dataBytes bytes of cells (the sizes come from
// tq_join_next_bytes) and returns the writable view.
func (c *Column) PrepareVarLenResult(n int, dataBytes int64, v *CColumn) {
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
	c.CView(v)
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
	c := &Column{elemBuf: make([]byte, elemLen), nullBitmap: make([]byte, 0, (capRows+7)>>3)}
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
