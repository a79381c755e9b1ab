"""Standalone deployment: Master / Worker daemons, executor backend, application client.

Names are resolved lazily so that ``python -m asyncframework_b200.deploy.master`` (and ``.worker`` / ``.executor``) do not find
their own module pre-imported by the package."""
_EXPORTS = {"Master": "master", "start_master": "master", "Worker": "worker", "start_worker": "worker",
            "LocalStandaloneCluster": "client", "StandaloneTransport": "client", "run_fused_on_cluster": "fused_launcher"}


def __getattr__(name):
    mod = _EXPORTS.get(name)
    if mod is None:
        raise AttributeError(name)
    import importlib
    return getattr(importlib.import_module(f"{__name__}.{mod}"), name)


__all__ = sorted(_EXPORTS)
