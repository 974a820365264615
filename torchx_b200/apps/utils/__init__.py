"""Utility programs that run as (or around) a role's entrypoint."""
