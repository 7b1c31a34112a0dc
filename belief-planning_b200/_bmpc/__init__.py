"""Internal host-side plumbing of the B200 Branch-MPC library (ctypes ABI mirror, config translation, batched solver)."""
