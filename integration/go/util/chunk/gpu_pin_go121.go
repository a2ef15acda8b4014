// +build gpu,go1.21

package chunk

import "runtime"

// pinner wraps runtime.Pinner (Go >= 1.21): a pinned Go buffer may be referenced from C memory for the duration of a call.
type pinner struct{ p runtime.Pinner }

func (pinner) available() bool        { return true }
func (x *pinner) Pin(ptr interface{}) { x.p.Pin(ptr) }
func (x *pinner) Unpin()              { x.p.Unpin() }
