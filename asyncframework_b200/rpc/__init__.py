"""Networked control plane (NettyRpcEnv / network-common analogue): see :mod:`.env`."""
from .env import (RpcAddress, RpcCallContext, RpcDisconnected, RpcEndpoint, RpcEndpointNotFound, RpcEndpointRef,  # noqa: F401
                  RpcEnv, RpcError, RpcRemoteFailure, RpcTimeout, parse_master_urls)
