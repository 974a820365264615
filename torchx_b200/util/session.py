"""One random session id per launcher process tree, inherited by children through the environment
(reference torchx/util/session.py:21-44)."""
import os
import uuid

from torchx_b200.settings import TORCHX_INTERNAL_SESSION_ID

_CURRENT = None


def get_session_id_or_create_new() -> str:
    global _CURRENT
    if _CURRENT:
        return _CURRENT
    _CURRENT = os.environ.get(TORCHX_INTERNAL_SESSION_ID) or str(uuid.uuid4())
    return _CURRENT
