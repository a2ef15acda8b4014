// +build gpu

// FinalMode HashAggExec over pushed-down partial results (SURVEY §8 f4).  When the planner pushes the Partial1 half of an
// aggregation into the coprocessor (planner/core/task.go:564-625 BuildFinalModeAggregation), the HashAggExec left in the root
// task carries AggFuncDescs with Mode == FinalMode whose Args are columns of the partial schema: per function its
// GetPartialResult columns (AVG: count, sum), then the GROUP BY columns (store/mockstore/mocktikv/aggregate.go:98-108).
// executorBuilder.buildHashAgg sends such a plan here instead of to GPUHashAggExec.Open's tq_agg_create.
package executor

/*
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
	"github.com/pingcap/tidb/util/chunk"
)

// isFinalMode: every function consumes partial data (aggregation.go:88-98; descriptor.go:52-75 builds them as a set).
func isFinalMode(funcs []*aggregation.AggFuncDesc) bool {
	for _, f := range funcs {
		if f.Mode != aggregation.FinalMode && f.Mode != aggregation.Partial2Mode {
			return false
		}
	}
	return len(funcs) > 0
}

// openFinal replaces the tq_agg_create call of GPUHashAggExec.Open; Next / Close are unchanged (tq_agg_put takes the
// child's chunks of partial rows, tq_agg_next returns final values).
func (e *GPUHashAggExec) openFinal(ctx context.Context) error {
	childTypes := retTypes(e.children[0])
	it := cInt32s(len(childTypes), func(i int) C.int32_t { return tqType(childTypes[i]) })
	gb := cInt32s(len(e.groupByItems), func(i int) C.int32_t { return C.int32_t(e.groupByItems[i].(*expression.Column).Index) })
	funcs := (*[1 << 8]C.tq_agg_final_func)(C.calloc(C.size_t(len(e.aggFuncs)), C.sizeof_tq_agg_final_func))
	d := (*C.tq_agg_final_desc)(C.calloc(1, C.sizeof_tq_agg_final_desc))
	defer func() {
		for _, p := range []unsafe.Pointer{unsafe.Pointer(it), unsafe.Pointer(gb), unsafe.Pointer(funcs), unsafe.Pointer(d)} {
			C.free(p)
		}
	}()
	for i, f := range e.aggFuncs {
		kind, _ := aggKind(f.Name)
		funcs[i]._func = kind
		funcs[i].arg_col = C.int32_t(f.Args[0].(*expression.Column).Index) // AVG: args[0] = partial count (func_avg.go:97)
		funcs[i].arg_col2 = -1
		if f.Name == ast.AggFuncAvg {
			funcs[i].arg_col2 = C.int32_t(f.Args[1].(*expression.Column).Index) // args[1] = partial sum (func_avg.go:89)
		}
	}
	d.n_input_cols, d.input_types = C.int32_t(len(childTypes)), it
	d.n_group_by, d.group_by_cols = C.int32_t(len(e.groupByItems)), gb
	d.n_funcs, d.funcs = C.int32_t(len(e.aggFuncs)), &funcs[0]
	d.est_groups = C.int64_t(e.estGroups)
	if st := C.tq_agg_create_final(d, &e.h); st != C.TQ_OK {
		return chunk.StatusError(int32(st))
	}
	return nil
}
