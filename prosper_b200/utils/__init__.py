"""Host-side helpers mirroring prosper.utils for the hot path (parallel, datalog)."""
