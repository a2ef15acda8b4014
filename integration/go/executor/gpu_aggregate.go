// +build gpu

// GPUHashAggExec: HashAggExec + its partial / final workers (executor/aggregate.go:54-155,482-588) replaced by
// libtinysql_b200.  executorBuilder.buildHashAgg (builder.go:486-542) returns it when every GROUP BY item and aggregate
// argument is a column (or has been pre-projected into one) of an 8-byte type.
package executor

/*
#cgo CFLAGS: -I${SRCDIR}/../../../include
#cgo LDFLAGS: -ltinysql_b200
#include <stdlib.h>
#include "tinysql_b200.h"
*/
import "C"

import (
	"context"
	"unsafe"

	"github.com/pingcap/tidb/expression"
	"github.com/pingcap/tidb/expression/aggregation"
	"github.com/pingcap/tidb/parser/ast"
	"github.com/pingcap/tidb/parser/mysql"
	"github.com/pingcap/tidb/types"
	"github.com/pingcap/tidb/util/chunk"
)

type GPUHashAggExec struct {
	baseExecutor

	aggFuncs     []*aggregation.AggFuncDesc
	groupByItems []expression.Expression
	estGroups    int64

	h           *C.tq_agg
	childResult *chunk.Chunk
	inViews     *chunk.CViewSet // argument block in C memory (cgo pointer rules, util/chunk/gpu_bridge.go)
	pump        resultPump      // result chunks incl. FLOAT / var-len columns (gpu_sort_merge.go)
	prepared    bool
}

func aggKind(name string) (C.int32_t, bool) {
	switch name {
	case ast.AggFuncCount:
		return C.TQ_AGG_COUNT, true
	case ast.AggFuncSum:
		return C.TQ_AGG_SUM, true
	case ast.AggFuncAvg:
		return C.TQ_AGG_AVG, true
	case ast.AggFuncMax:
		return C.TQ_AGG_MAX, true
	case ast.AggFuncMin:
		return C.TQ_AGG_MIN, true
	case ast.AggFuncFirstRow:
		return C.TQ_AGG_FIRSTROW, true
	}
	return 0, false
}

// Open implements Executor (aggregate.go:199-223).
func (e *GPUHashAggExec) Open(ctx context.Context) error {
	if err := e.baseExecutor.Open(ctx); err != nil {
		return err
	}
	childTypes := retTypes(e.children[0])
	it := make([]C.int32_t, len(childTypes))
	for i := range it {
		it[i] = tqType(childTypes[i])
		if mysql.HasNotNullFlag(childTypes[i].Flag) {
			it[i] |= C.TQ_TYPE_NOT_NULL // lets SUM / MAX / MIN drop their "saw a value" word
		}
	}
	gb := make([]C.int32_t, len(e.groupByItems))
	for i := range gb {
		gb[i] = C.int32_t(e.groupByItems[i].(*expression.Column).Index)
	}
	funcs := make([]C.tq_agg_func, len(e.aggFuncs))
	for i, f := range e.aggFuncs {
		kind, _ := aggKind(f.Name)
		funcs[i]._func = kind // cgo renames the C field `func`
		funcs[i].arg_col = -1 // a constant non-NULL argument: COUNT(*) == count(1) (parser.y:3258-3262)
		if col, ok := f.Args[0].(*expression.Column); ok {
			funcs[i].arg_col = C.int32_t(col.Index)
		}
	}
	if isFinalMode(e.aggFuncs) { // pushed-down partial results: gpu_aggregate_final.go
		if err := e.openFinal(ctx); err != nil {
			return err
		}
	} else {
		// the descriptor and its arrays live in C memory for the duration of the call
		d := (*C.tq_agg_desc)(C.calloc(1, C.sizeof_tq_agg_desc))
		cit := cInt32s(len(it), func(i int) C.int32_t { return it[i] })
		cgb := cInt32s(len(gb), func(i int) C.int32_t { return gb[i] })
		cf := (*[1 << 8]C.tq_agg_func)(C.calloc(C.size_t(len(funcs)), C.sizeof_tq_agg_func))
		copy(cf[:len(funcs)], funcs)
		d.n_input_cols, d.input_types = C.int32_t(len(it)), cit
		d.n_group_by, d.group_by_cols = C.int32_t(len(gb)), cgb
		d.n_funcs, d.funcs = C.int32_t(len(funcs)), &cf[0]
		d.est_groups = C.int64_t(e.estGroups) // the planner's NDV estimate sizes the table and enables pre-aggregation
		st := C.tq_agg_create(d, &e.h)
		for _, p := range []unsafe.Pointer{unsafe.Pointer(d), unsafe.Pointer(cit), unsafe.Pointer(cgb), unsafe.Pointer(cf)} {
			C.free(p)
		}
		if st != C.TQ_OK {
			return chunk.StatusError(int32(st))
		}
	}
	e.childResult = newFirstChunk(e.children[0])
	e.inViews = chunk.NewCViewSet(len(childTypes))
	e.pump = resultPump{outViews: chunk.NewCViewSet(len(e.aggFuncs)), sizes: (*C.int64_t)(C.calloc(C.size_t(len(e.aggFuncs)), 8))}
	for i := range e.aggFuncs {
		var t C.int32_t
		C.tq_agg_output_type(e.h, C.int32_t(i), &t) // COUNT -> BIGINT; MAX / MIN / FIRSTROW keep FLOAT / string columns
		e.pump.outTypes = append(e.pump.outTypes, t)
	}
	e.prepared = false
	return nil
}

// Next implements Executor (aggregate.go:482-588): a pipeline breaker — the first call drains the child.
func (e *GPUHashAggExec) Next(ctx context.Context, req *chunk.Chunk) error {
	req.Reset()
	if !e.prepared {
		for { // fetchChildData (aggregate.go:487-522)
			if err := Next(ctx, e.children[0], e.childResult); err != nil {
				return err
			}
			if e.childResult.NumRows() == 0 {
				break
			}
			e.inViews.FillChunk(e.childResult)
			st := C.tq_agg_put(e.h, e.inViews.Ptr(), C.TQ_MEM_HOST)
			e.inViews.Release()
			if st != C.TQ_OK {
				return chunk.StatusError(int32(st))
			}
		}
		if st := C.tq_agg_eof(e.h); st != C.TQ_OK {
			return chunk.StatusError(int32(st))
		}
		e.prepared = true
	}
	// the default row of an empty scalar aggregate (aggregate.go:572-574) is produced by the library;
	// SUM(BIGINT) overflow surfaces here as types.ErrOverflow (func_sum.go:133-136)
	err := e.pump.fill(req,
		func(want C.int64_t, sizes *C.int64_t) C.int32_t { return C.tq_agg_next_bytes(e.h, want, sizes) },
		func(want C.int64_t, out *C.tq_column, n *C.int64_t, eof *C.int32_t) C.int32_t { return C.tq_agg_next(e.h, want, out, n, eof) })
	if se, ok := err.(chunk.StatusError); ok {
		return statusToAggError(C.int32_t(se))
	}
	return err
}

// statusToAggError: SUM / AVG over BIGINT report types.ErrOverflow like types.AddInt64 does (func_sum.go:133-136).
func statusToAggError(st C.int32_t) error {
	if st == C.TQ_ERR_OVERFLOW_BIGINT {
		return types.ErrOverflow.GenWithStackByArgs("BIGINT", "sum")
	}
	return chunk.StatusError(int32(st))
}

// Close implements Executor; Close may run after Open without Next (aggregate.go:187-197).
func (e *GPUHashAggExec) Close() error {
	if e.h != nil {
		C.tq_agg_destroy(e.h)
		e.h = nil
		e.inViews.Free()
		e.pump.outViews.Free()
		C.free(unsafe.Pointer(e.pump.sizes))
	}
	return e.baseExecutor.Close()
}
