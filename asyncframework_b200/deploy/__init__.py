"""Standalone deployment: Master / Worker daemons, executor backend, application client."""
from .master import Master, start_master  # noqa: F401
from .worker import Worker, start_worker  # noqa: F401
from .client import LocalStandaloneCluster, StandaloneTransport  # noqa: F401
from .fused_launcher import run_fused_on_cluster  # noqa: F401
