// +build gpu,!go1.21

package chunk

// Toolchains without runtime.Pinner: CViewSet copies Go buffers into C scratch instead (strategy c in gpu_bridge.go).
type pinner struct{}

func (pinner) available() bool     { return false }
func (*pinner) Pin(ptr interface{}) {}
func (*pinner) Unpin()             {}
