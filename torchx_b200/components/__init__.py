"""Builtin components (functions returning an AppDef).  ``dist.ddp`` is the one the hot path is launched through."""
