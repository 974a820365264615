"""One random session id per launcher process tree, inherited by children through the environment
(reference torchx/util/session.py:21-44)."""
import os
import uuid

from torchx_b200.settings import TORCHX_INTERNAL_SESSION_ID

CURRENT_SESSION_ID = None  # module-level so tests and embedding code can reset it


def get_session_id_or_create_new() -> str:
    global CURRENT_SESSION_ID
    if CURRENT_SESSION_ID:
        return CURRENT_SESSION_ID
    CURRENT_SESSION_ID = os.environ.get(TORCHX_INTERNAL_SESSION_ID) or str(uuid.uuid4())
    return CURRENT_SESSION_ID


def get_torchx_session_id():
    """The session id if one has been created or inherited, else None (for code outside the launcher)."""
    return CURRENT_SESSION_ID
