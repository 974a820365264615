"""Worker-side data plane: communicator, DDP wrapper and comm hook on libb200ddp.so."""
from .comm import Communicator, local_pass_, mode_for  # noqa: F401
