// +build gpu

// GPUHashJoinExec: HashJoinExec (executor/join.go:31-146) with the hash table, the join workers and the joiners replaced
// by libtinysql_b200.  executorBuilder.buildHashJoin (builder.go:431-484) returns newGPUHashJoin(...) instead of
// &HashJoinExec{...} when every join key is an 8-byte type; otherwise it keeps the Go executor.
package executor

/*
#cgo CFLAGS: -I${SRCDIR}/../../../include
#cgo LDFLAGS: -ltinysql_b200
#include "tinysql_b200.h"
*/
import "C"

import (
	"context"
	"unsafe"

	"github.com/pingcap/tidb/expression"
	"github.com/pingcap/tidb/parser/mysql"
	plannercore "github.com/pingcap/tidb/planner/core"
	"github.com/pingcap/tidb/types"
	"github.com/pingcap/tidb/util/chunk"
)

// tqType maps a FieldType onto the C-ABI's column types (util/chunk/codec.go:171-181 getFixedLen decides the layout).
func tqType(ft *types.FieldType) C.int32_t {
	switch ft.Tp {
	case mysql.TypeFloat:
		return C.TQ_TYPE_FLOAT32
	case mysql.TypeDouble:
		return C.TQ_TYPE_FLOAT64
	case mysql.TypeTiny, mysql.TypeShort, mysql.TypeInt24, mysql.TypeLong, mysql.TypeLonglong, mysql.TypeYear:
		if mysql.HasUnsignedFlag(ft.Flag) {
			return C.TQ_TYPE_UINT64
		}
		return C.TQ_TYPE_INT64
	default:
		return C.TQ_TYPE_BYTES
	}
}

type GPUHashJoinExec struct {
	baseExecutor

	outerSideExec, innerSideExec Executor
	outerSideFilter              expression.CNFExprs
	outerKeys, innerKeys         []*expression.Column
	otherConditions              expression.CNFExprs // applied on the returned chunk (inner joins), see Next
	joinType                     plannercore.JoinType
	outerIsRight                 bool

	h         *C.tq_join
	innerChk  *chunk.Chunk
	outerChk  *chunk.Chunk
	selected  []bool
	selBytes  []byte
	outTypes  []C.int32_t
	views     []chunk.CColumn
	sizes     []C.int64_t
	prepared  bool
	outerDone bool
}

func i32s(n int, f func(i int) C.int32_t) []C.int32_t {
	s := make([]C.int32_t, n)
	for i := range s {
		s[i] = f(i)
	}
	return s
}

// Open implements Executor (join.go:110-123).
func (e *GPUHashJoinExec) Open(ctx context.Context) error {
	if err := e.baseExecutor.Open(ctx); err != nil {
		return err
	}
	innerTypes, outerTypes := retTypes(e.innerSideExec), retTypes(e.outerSideExec)
	bt := i32s(len(innerTypes), func(i int) C.int32_t { return tqType(innerTypes[i]) })
	pt := i32s(len(outerTypes), func(i int) C.int32_t { return tqType(outerTypes[i]) })
	bk := i32s(len(e.innerKeys), func(i int) C.int32_t { return C.int32_t(e.innerKeys[i].Index) })
	pk := i32s(len(e.outerKeys), func(i int) C.int32_t { return C.int32_t(e.outerKeys[i].Index) })
	var d C.tq_join_desc
	d.join_type = C.int32_t(e.joinType) // InnerJoin=0, LeftOuterJoin=1, RightOuterJoin=2 (logical_plans.go:52-57)
	if e.outerIsRight {
		d.outer_is_right = 1
	}
	d.n_build_cols, d.build_types = C.int32_t(len(bt)), &bt[0]
	d.n_probe_cols, d.probe_types = C.int32_t(len(pt)), &pt[0]
	d.n_keys, d.build_key_idx, d.probe_key_idx = C.int32_t(len(bk)), &bk[0], &pk[0]
	if st := C.tq_join_create(&d, &e.h); st != C.TQ_OK {
		return chunk.StatusError(int32(st))
	}
	// output = lhs ++ rhs (joiner.go:145-150)
	lhs, rhs := pt, bt
	if e.outerIsRight {
		lhs, rhs = bt, pt
	}
	e.outTypes = append(append([]C.int32_t{}, lhs...), rhs...)
	e.views = make([]chunk.CColumn, len(e.outTypes))
	e.sizes = make([]C.int64_t, len(e.outTypes))
	e.innerChk, e.outerChk = newFirstChunk(e.innerSideExec), newFirstChunk(e.outerSideExec)
	e.prepared, e.outerDone = false, false
	return nil
}

// fetchAndBuildHashTable (join.go:148-158): drain the inner child into the device-side row container.
func (e *GPUHashJoinExec) build(ctx context.Context) error {
	views := make([]chunk.CColumn, e.innerChk.NumCols())
	for {
		if err := Next(ctx, e.innerSideExec, e.innerChk); err != nil {
			return err
		}
		if e.innerChk.NumRows() == 0 {
			break
		}
		e.innerChk.CViews(views)
		if st := C.tq_join_put_build(e.h, &views[0], C.TQ_MEM_HOST); st != C.TQ_OK {
			return chunk.StatusError(int32(st))
		}
	}
	if st := C.tq_join_finalize_build(e.h); st != C.TQ_OK {
		return chunk.StatusError(int32(st))
	}
	return nil
}

// Next implements Executor (join.go:125-146): fills req with <= req.RequiredRows() joined rows; 0 rows == EOF.
func (e *GPUHashJoinExec) Next(ctx context.Context, req *chunk.Chunk) error {
	req.Reset()
	if !e.prepared {
		if err := e.build(ctx); err != nil {
			return err
		}
		e.prepared = true
	}
	want := req.RequiredRows()
	outerViews := make([]chunk.CColumn, e.outerChk.NumCols())
	for {
		// size the result columns for this call (the *_next_size query of the ownership contract)
		if st := C.tq_join_next_bytes(e.h, C.int64_t(want), &e.sizes[0]); st != C.TQ_OK {
			return chunk.StatusError(int32(st))
		}
		for i := range e.views {
			col := req.Column(i)
			switch e.outTypes[i] {
			case C.TQ_TYPE_BYTES:
				col.PrepareVarLenResult(want, int64(e.sizes[i]), &e.views[i])
			case C.TQ_TYPE_FLOAT32:
				col.PrepareFixedResult(want, 4, &e.views[i])
			default:
				col.PrepareFixedResult(want, 8, &e.views[i])
			}
		}
		var n C.int64_t
		var eof C.int32_t
		if st := C.tq_join_next(e.h, C.int64_t(want), &e.views[0], &n, &eof); st != C.TQ_OK {
			return chunk.StatusError(int32(st))
		}
		if n > 0 || eof != 0 {
			for i := range e.views {
				req.Column(i).SetResultRows(int(n))
			}
			req.SetNumVirtualRows(int(n))
			if n > 0 && len(e.otherConditions) > 0 {
				// inner joins: baseJoiner.filter (joiner.go:155-167) on the returned chunk; outer joins with
				// OtherConditions use tq_join_set_other_conditions (comparison conditions) or stay on the Go executor
				var err error
				if e.selected, err = expression.VectorizedFilter(e.ctx, e.otherConditions, chunk.NewIterator4Chunk(req), e.selected); err != nil {
					return err
				}
				req.SetSel(selToIdx(e.selected))
				if req.NumRows() == 0 && eof == 0 {
					req.Reset()
					continue
				}
			}
			return nil
		}
		// fetchOuterSideChunks (join.go:194-221): feed one more outer chunk
		if e.outerDone {
			continue
		}
		if err := Next(ctx, e.outerSideExec, e.outerChk); err != nil {
			return err
		}
		if e.outerChk.NumRows() == 0 {
			e.outerDone = true
			C.tq_join_probe_eof(e.h)
			continue
		}
		var sel *C.uint8_t
		if len(e.outerSideFilter) > 0 { // join.go:328: rows the outer-side filter rejects are misses
			var err error
			if e.selected, err = expression.VectorizedFilter(e.ctx, e.outerSideFilter, chunk.NewIterator4Chunk(e.outerChk), e.selected); err != nil {
				return err
			}
			e.selBytes = e.selBytes[:0]
			for _, s := range e.selected {
				if s {
					e.selBytes = append(e.selBytes, 1)
				} else {
					e.selBytes = append(e.selBytes, 0)
				}
			}
			sel = (*C.uint8_t)(unsafe.Pointer(&e.selBytes[0]))
		}
		e.outerChk.CViews(outerViews)
		if st := C.tq_join_put_probe(e.h, &outerViews[0], sel, C.TQ_MEM_HOST); st != C.TQ_OK {
			return chunk.StatusError(int32(st))
		}
	}
}

func selToIdx(selected []bool) []int {
	idx := make([]int, 0, len(selected))
	for i, s := range selected {
		if s {
			idx = append(idx, i)
		}
	}
	return idx
}

// Close implements Executor (join.go:81-108); safe straight after Open and with rows still pending (`limit 1`).
func (e *GPUHashJoinExec) Close() error {
	if e.h != nil {
		C.tq_join_destroy(e.h)
		e.h = nil
	}
	return e.baseExecutor.Close()
}
