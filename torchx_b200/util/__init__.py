"""Small helpers shared by the launcher modules (types, session ids, entry points, log tee)."""
