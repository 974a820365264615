"""Programs shipped with the launcher (see ``apps.utils``)."""
