// +build gpu

// GPU-backed vectorized builtin signatures.  The function classes (compareFunctionClass.getFunction
// builtin_compare.go:102, arithmetic*FunctionClass.getFunction builtin_arithmetic.go:112-352, ...) wrap the signature they
// would have returned:  sig = &gpuCompareIntSig{builtinLTIntSig: sig, op: C.TQ_CMP_LT}.  Argument evaluation stays in Go
// (b.args[i].VecEvalInt — builtin_compare_vec.go:186-205); only the element-wise loop moves to the device.
package expression

/*
#cgo CFLAGS: -I${SRCDIR}/../../../include
#cgo LDFLAGS: -ltinysql_b200
#include "tinysql_b200.h"
*/
import "C"

import (
	"unsafe"

	"github.com/pingcap/tidb/parser/mysql"
	"github.com/pingcap/tidb/types"
	"github.com/pingcap/tidb/util/chunk"
)

func cbool(b bool) C.int32_t {
	if b {
		return 1
	}
	return 0
}

// statusToError maps the library's status onto the reference's error values (types.ErrOverflow with the same
// arguments the Go loops use, builtin_arithmetic_vec.go:52,441,489).
func statusToError(st C.int32_t, expr string) error {
	switch st {
	case C.TQ_OK:
		return nil
	case C.TQ_ERR_OVERFLOW_BIGINT:
		return types.ErrOverflow.GenWithStackByArgs("BIGINT", expr)
	case C.TQ_ERR_OVERFLOW_BIGINT_UNSIGNED:
		return types.ErrOverflow.GenWithStackByArgs("BIGINT UNSIGNED", expr)
	case C.TQ_ERR_OVERFLOW_DOUBLE:
		return types.ErrOverflow.GenWithStackByArgs("DOUBLE", expr)
	default:
		return chunk.StatusError(int32(st))
	}
}

// evalTwoIntArgs is the prologue every Go signature has (builtin_compare_vec.go:186-205).
func evalTwoArgs(b *baseBuiltinFunc, input *chunk.Chunk, et types.EvalType) (buf0, buf1 *chunk.Column, release func(), err error) {
	n := input.NumRows()
	if buf0, err = b.bufAllocator.get(et, n); err != nil {
		return
	}
	if buf1, err = b.bufAllocator.get(et, n); err != nil {
		b.bufAllocator.put(buf0)
		return
	}
	release = func() { b.bufAllocator.put(buf0); b.bufAllocator.put(buf1) }
	eval := func(i int, buf *chunk.Column) error {
		switch et {
		case types.ETInt:
			return b.args[i].VecEvalInt(b.ctx, input, buf)
		case types.ETReal:
			return b.args[i].VecEvalReal(b.ctx, input, buf)
		default:
			return b.args[i].VecEvalString(b.ctx, input, buf)
		}
	}
	if err = eval(0, buf0); err == nil {
		err = eval(1, buf1)
	}
	if err != nil {
		release()
	}
	return
}

// twoArgCall is the per-signature argument block of a tq_vec_* call: three tq_column structs in C memory (cgo pointer rules:
// a tq_column in Go memory would hold Go pointers — util/chunk/gpu_bridge.go) — entries 0 / 1 the operands, 2 the result.
type twoArgCall struct{ v *chunk.CViewSet }

func (c *twoArgCall) begin(buf0, buf1, result *chunk.Column, n, elemLen int) (a, b, out *C.tq_column) {
	if c.v == nil {
		c.v = chunk.NewCViewSet(3)
	}
	c.v.Fill(0, buf0)
	c.v.Fill(1, buf1)
	result.PrepareFixedResult(n, elemLen)
	c.v.FillResult(2, result)
	return (*C.tq_column)(unsafe.Pointer(c.v.At(0))), (*C.tq_column)(unsafe.Pointer(c.v.At(1))), (*C.tq_column)(unsafe.Pointer(c.v.At(2)))
}

// end unpins the operands and, on success, moves the produced rows into the result column (a no-op when the result buffer was
// pinned or C-owned and written in place).
func (c *twoArgCall) end(result *chunk.Column, n int, st C.int32_t) {
	c.v.Release()
	if st == C.TQ_OK {
		c.v.CopyBack(2, result, n)
		result.SetResultRows(n)
	}
}

// gpuCompareIntSig: builtin{LT,LE,GT,GE,EQ,NE}IntSig.vecEvalInt (builtin_compare_vec.go:22-292)
type gpuCompareIntSig struct {
	baseBuiltinFunc
	op   C.int32_t
	call twoArgCall
}

func (b *gpuCompareIntSig) vectorized() bool { return true }

func (b *gpuCompareIntSig) vecEvalInt(input *chunk.Chunk, result *chunk.Column) error {
	n := input.NumRows()
	buf0, buf1, release, err := evalTwoArgs(&b.baseBuiltinFunc, input, types.ETInt)
	if err != nil {
		return err
	}
	defer release()
	a, bb, out := b.call.begin(buf0, buf1, result, n, 8)
	st := C.tq_vec_compare_int(b.op, C.int64_t(n),
		a, cbool(mysql.HasUnsignedFlag(b.args[0].GetType().Flag)),
		bb, cbool(mysql.HasUnsignedFlag(b.args[1].GetType().Flag)), out, C.TQ_MEM_HOST)
	b.call.end(result, n, st)
	return statusToError(st, "")
}

// gpuCompareRealSig / gpuCompareStringSig: builtin{LT..NE}{Real,String}Sig (builtin_compare_vec_generated.go)
type gpuCompareRealSig struct {
	baseBuiltinFunc
	op   C.int32_t
	call twoArgCall
}

func (b *gpuCompareRealSig) vectorized() bool { return true }

func (b *gpuCompareRealSig) vecEvalInt(input *chunk.Chunk, result *chunk.Column) error {
	n := input.NumRows()
	buf0, buf1, release, err := evalTwoArgs(&b.baseBuiltinFunc, input, types.ETReal)
	if err != nil {
		return err
	}
	defer release()
	a, bb, out := b.call.begin(buf0, buf1, result, n, 8)
	st := C.tq_vec_compare_real(b.op, C.int64_t(n), a, bb, out, C.TQ_MEM_HOST)
	b.call.end(result, n, st)
	return statusToError(st, "")
}

type gpuCompareStringSig struct {
	baseBuiltinFunc
	op   C.int32_t // TQ_CMP_* or TQ_STR_STRCMP (builtinStrcmpSig, builtin_string_vec.go:52-83)
	call twoArgCall
}

func (b *gpuCompareStringSig) vectorized() bool { return true }

func (b *gpuCompareStringSig) vecEvalInt(input *chunk.Chunk, result *chunk.Column) error {
	n := input.NumRows()
	buf0, buf1, release, err := evalTwoArgs(&b.baseBuiltinFunc, input, types.ETString)
	if err != nil {
		return err
	}
	defer release()
	a, bb, out := b.call.begin(buf0, buf1, result, n, 8)
	st := C.tq_vec_compare_string(b.op, C.int64_t(n), a, bb, out, C.TQ_MEM_HOST)
	b.call.end(result, n, st)
	return statusToError(st, "")
}

// gpuArithIntSig: builtinArithmetic{Plus,Minus,Multiply}IntSig and MultiplyIntUnsignedSig
// (builtin_arithmetic_vec.go:88-340,389-532)
type gpuArithIntSig struct {
	baseBuiltinFunc
	op   C.int32_t
	expr string // "(%s + %s)" text for the overflow error, as the Go loops build it
	call twoArgCall
}

func (b *gpuArithIntSig) vectorized() bool { return true }

func (b *gpuArithIntSig) vecEvalInt(input *chunk.Chunk, result *chunk.Column) error {
	n := input.NumRows()
	buf0, buf1, release, err := evalTwoArgs(&b.baseBuiltinFunc, input, types.ETInt)
	if err != nil {
		return err
	}
	defer release()
	a, bb, out := b.call.begin(buf0, buf1, result, n, 8)
	st := C.tq_vec_arith_int(b.op, C.int64_t(n),
		a, cbool(mysql.HasUnsignedFlag(b.args[0].GetType().Flag)),
		bb, cbool(mysql.HasUnsignedFlag(b.args[1].GetType().Flag)), out, C.TQ_MEM_HOST)
	b.call.end(result, n, st)
	return statusToError(st, b.expr)
}

// gpuArithRealSig: builtinArithmetic{Plus,Minus,Multiply,Divide}RealSig (builtin_arithmetic_vec.go:25-86,282-387)
type gpuArithRealSig struct {
	baseBuiltinFunc
	op   C.int32_t
	expr string
	call twoArgCall
}

func (b *gpuArithRealSig) vectorized() bool { return true }

func (b *gpuArithRealSig) vecEvalReal(input *chunk.Chunk, result *chunk.Column) error {
	n := input.NumRows()
	buf0, buf1, release, err := evalTwoArgs(&b.baseBuiltinFunc, input, types.ETReal)
	if err != nil {
		return err
	}
	defer release()
	a, bb, out := b.call.begin(buf0, buf1, result, n, 8)
	var divByZero C.int64_t // a Go int64 holds no pointers: legal as a direct argument
	st := C.tq_vec_arith_real(b.op, C.int64_t(n), a, bb, out, &divByZero, C.TQ_MEM_HOST)
	b.call.end(result, n, st)
	if err := statusToError(st, b.expr); err != nil {
		return err
	}
	for i := C.int64_t(0); i < divByZero; i++ { // one warning (or error, by SQL mode) per NULL-ed row: :369-375
		if err := handleDivisionByZeroError(b.ctx); err != nil {
			return err
		}
	}
	return nil
}
