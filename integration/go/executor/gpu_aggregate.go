// +build gpu

// GPUHashAggExec: HashAggExec + its partial / final workers (executor/aggregate.go:54-155,482-588) replaced by
// libtinysql_b200.  executorBuilder.buildHashAgg (builder.go:486-542) returns it when every GROUP BY item and aggregate
// argument is a column (or has been pre-projected into one) of an 8-byte type.
package executor

/*
#cgo CFLAGS: -I${SRCDIR}/../../../include
#cgo LDFLAGS: -ltinysql_b200
#include "tinysql_b200.h"
*/
import "C"

import (
	"context"

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
	views       []chunk.CColumn
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
	it := i32s(len(childTypes), func(i int) C.int32_t {
		t := tqType(childTypes[i])
		if mysql.HasNotNullFlag(childTypes[i].Flag) {
			t |= C.TQ_TYPE_NOT_NULL // lets SUM / MAX / MIN drop their "saw a value" word
		}
		return t
	})
	gb := i32s(len(e.groupByItems), func(i int) C.int32_t { return C.int32_t(e.groupByItems[i].(*expression.Column).Index) })
	funcs := make([]C.tq_agg_func, len(e.aggFuncs))
	for i, f := range e.aggFuncs {
		kind, _ := aggKind(f.Name)
		funcs[i].func = kind
		funcs[i].arg_col = -1 // a constant non-NULL argument: COUNT(*) == count(1) (parser.y:3258-3262)
		if col, ok := f.Args[0].(*expression.Column); ok {
			funcs[i].arg_col = C.int32_t(col.Index)
		}
	}
	var d C.tq_agg_desc
	d.n_input_cols, d.input_types = C.int32_t(len(it)), &it[0]
	d.n_group_by = C.int32_t(len(gb))
	if len(gb) > 0 {
		d.group_by_cols = &gb[0]
	}
	d.n_funcs, d.funcs = C.int32_t(len(funcs)), &funcs[0]
	d.est_groups = C.int64_t(e.estGroups) // the planner's NDV estimate sizes the table and enables pre-aggregation
	if st := C.tq_agg_create(&d, &e.h); st != C.TQ_OK {
		return chunk.StatusError(int32(st))
	}
	e.childResult = newFirstChunk(e.children[0])
	e.views = make([]chunk.CColumn, len(funcs))
	e.prepared = false
	return nil
}

// Next implements Executor (aggregate.go:482-588): a pipeline breaker — the first call drains the child.
func (e *GPUHashAggExec) Next(ctx context.Context, req *chunk.Chunk) error {
	req.Reset()
	if !e.prepared {
		in := make([]chunk.CColumn, e.childResult.NumCols())
		for { // fetchChildData (aggregate.go:487-522)
			if err := Next(ctx, e.children[0], e.childResult); err != nil {
				return err
			}
			if e.childResult.NumRows() == 0 {
				break
			}
			e.childResult.CViews(in)
			if st := C.tq_agg_put(e.h, &in[0], C.TQ_MEM_HOST); st != C.TQ_OK {
				return chunk.StatusError(int32(st))
			}
		}
		if st := C.tq_agg_eof(e.h); st != C.TQ_OK {
			return chunk.StatusError(int32(st))
		}
		e.prepared = true
	}
	want := req.RequiredRows()
	for i := range e.views {
		req.Column(i).PrepareFixedResult(want, 8, &e.views[i])
	}
	var n C.int64_t
	var eof C.int32_t
	// the default row of an empty scalar aggregate (aggregate.go:572-574) is produced by the library
	st := C.tq_agg_next(e.h, C.int64_t(want), &e.views[0], &n, &eof)
	if st != C.TQ_OK {
		return statusToAggError(st) // SUM(BIGINT) overflow -> types.ErrOverflow (func_sum.go:133-136)
	}
	for i := range e.views {
		req.Column(i).SetResultRows(int(n))
	}
	req.SetNumVirtualRows(int(n))
	return nil
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
	}
	return e.baseExecutor.Close()
}
