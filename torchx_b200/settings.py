"""Names of the ``TORCHX_*`` environment variables that cross the launcher -> worker boundary
(reference torchx/settings.py:22-37, runner/api.py:378-391).  Kept under the TorchX names on purpose: worker scripts
and trackers written for TorchX read exactly these."""
ENV_TORCHX_JOB_ID = "TORCHX_JOB_ID"
ENV_TORCHX_PARENT_RUN_ID = "TORCHX_PARENT_RUN_ID"
ENV_TORCHX_TRACKERS = "TORCHX_TRACKERS"
ENV_TORCHX_IMAGE = "TORCHX_IMAGE"
ENV_TORCHXCONFIG = "TORCHXCONFIG"
TORCHX_INTERNAL_SESSION_ID = "TORCHX_INTERNAL_SESSION_ID"
TORCHX_HOME = "TORCHX_HOME"


def tracker_config_env_var_name(tracker_name: str) -> str:
    """``TORCHX_TRACKER_<NAME>_CONFIG``: where a worker-side tracker backend finds its configuration
    (reference torchx/tracker/api.py:125-127)."""
    return f"TORCHX_TRACKER_{tracker_name.upper()}_CONFIG"
