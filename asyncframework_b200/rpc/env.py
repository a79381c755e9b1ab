"""TCP RPC environment: named endpoints, one-way ``send``, request/response ``ask``, raw chunk fetch.

The control plane of the reference is ``NettyRpcEnv`` + ``Dispatcher`` + per-endpoint ``Inbox`` + per-peer
``Outbox`` (core/rpc/netty/NettyRpcEnv.scala:43, Dispatcher.scala:39, Inbox.scala, Outbox.scala) over the
``common/network-common`` wire protocol (RpcRequest / RpcResponse / RpcFailure / OneWayMessage /
ChunkFetchRequest / ChunkFetchSuccess / ChunkFetchFailure frames, length-prefixed, decoded by
``TransportFrameDecoder``).  This module is the same capability for the multi-process / multi-node side of this
framework (standalone master/worker daemons, executor registration, task launch + status, block-manager master,
peer-to-peer block transfer).  The GPU hot path never touches it: gradients and parameters move over NVLink from
inside the kernels (``csrc/include/afb_round_impl.cuh``); this is for everything the reference does over Netty
that is *not* the per-round numeric traffic.

Design (not a transcription of Netty):
  * one TCP connection per (local env, remote address), bidirectional — replies and server-initiated messages
    ride the same socket, so an executor behind NAT-less localhost or a remote host needs only one dial-out;
  * frames: ``!IBQ`` = body length, frame type, request id; bodies of RPC frames are cloudpickle, bodies of chunk
    frames are raw bytes (no serialisation on the bulk path);
  * a HELLO handshake advertises the dialer's own listening address (so the callee can ``send`` back and so
    ``on_disconnected`` reports a meaningful peer) and optionally proves knowledge of a shared secret
    (HMAC-SHA256 over a server nonce — the role of ``spark.authenticate`` / SASL, common/network-common sasl/) and can
    switch the connection to AES-256-GCM (``encrypt=True``; ``spark.network.crypto.enabled``, network-common crypto/);
  * every endpoint owns an inbox thread: messages to one endpoint are processed sequentially, different endpoints
    run concurrently (``ThreadSafeRpcEndpoint`` semantics);
  * connection loss fails all outstanding asks on that connection with :class:`RpcDisconnected` and posts
    ``on_disconnected(address)`` to every endpoint — this is how the driver learns about a lost executor and how
    the master learns about a lost worker.
"""
from __future__ import annotations

import hashlib
import hmac
import itertools
import json
import os
import queue
import socket
import struct
import threading
import traceback
from concurrent.futures import Future
from typing import Any, Callable, Dict, NamedTuple, Optional, Tuple

import cloudpickle
import pickle

_HDR = struct.Struct("!IBQ")
F_HELLO, F_HELLO_ACK, F_ONEWAY, F_REQUEST, F_RESPONSE, F_FAILURE, F_CHUNK_REQ, F_CHUNK_OK, F_CHUNK_FAIL, F_CHALLENGE = range(10)
MAX_FRAME = 1 << 31


class RpcAddress(NamedTuple):
    host: str
    port: int

    def __str__(self) -> str:
        return f"{self.host}:{self.port}"

    @staticmethod
    def parse(url: str) -> "RpcAddress":
        """``host:port`` or ``spark://host:port`` / ``async://host:port``."""
        if "://" in url:
            url = url.split("://", 1)[1]
        host, _, port = url.rpartition(":")
        return RpcAddress(host, int(port))


def parse_master_urls(url: str) -> list:
    """``spark://h1:p1,h2:p2`` (the HA form of a standalone master URL) -> [RpcAddress, ...]."""
    body = url.split("://", 1)[1] if "://" in url else url
    return [RpcAddress.parse(part.strip()) for part in body.split(",") if part.strip()]


class RpcError(RuntimeError):
    pass


class RpcTimeout(RpcError):
    pass


class RpcDisconnected(RpcError):
    pass


class RpcEndpointNotFound(RpcError):
    pass


class RpcRemoteFailure(RpcError):
    """The remote endpoint raised; the message carries its repr and a short traceback."""


class RpcCallContext:
    """Handed to ``receive_and_reply``; the endpoint may reply later from another thread (like ``RpcCallContext``)."""

    def __init__(self, conn: "_Connection", req_id: int, sender: Optional[RpcAddress]):
        self._conn, self._req_id, self.sender_address = conn, req_id, sender
        self._done = False

    def reply(self, value: Any) -> None:
        if self._done:
            return
        self._done = True
        try:
            self._conn.send_frame(F_RESPONSE, self._req_id, cloudpickle.dumps(value))
        except OSError:
            pass

    def send_failure(self, exc: BaseException) -> None:
        if self._done:
            return
        self._done = True
        try:
            self._conn.send_frame(F_FAILURE, self._req_id, cloudpickle.dumps(f"{exc!r}\n{traceback.format_exc(limit=6)}"))
        except OSError:
            pass


class RpcEndpoint:
    """Override what you need.  ``self.rpc_env`` / ``self.self_ref`` are set by ``setup_endpoint``."""
    rpc_env: "RpcEnv" = None  # type: ignore[assignment]
    self_ref: "RpcEndpointRef" = None  # type: ignore[assignment]

    def on_start(self) -> None: ...
    def on_stop(self) -> None: ...
    def on_connected(self, address: RpcAddress) -> None: ...
    def on_disconnected(self, address: RpcAddress) -> None: ...

    def on_error(self, exc: BaseException) -> None:
        traceback.print_exception(type(exc), exc, exc.__traceback__)

    def receive(self, msg: Any, sender: Optional[RpcAddress]) -> None:
        raise RpcError(f"{type(self).__name__} does not handle one-way message {type(msg).__name__}")

    def receive_and_reply(self, msg: Any, ctx: RpcCallContext) -> None:
        ctx.send_failure(RpcError(f"{type(self).__name__} does not answer {type(msg).__name__}"))


class RpcEndpointRef:
    def __init__(self, env: "RpcEnv", address: Optional[RpcAddress], name: str):
        self._env, self.address, self.name = env, address, name

    def send(self, msg: Any) -> None:
        """Fire and forget (``RpcEndpointRef.send``).  Raises only if the peer cannot be reached at all."""
        self._env._send_oneway(self.address, self.name, msg)

    def ask(self, msg: Any) -> Future:
        return self._env._ask(self.address, self.name, msg)

    def ask_sync(self, msg: Any, timeout: Optional[float] = 30.0) -> Any:
        fut = self.ask(msg)
        try:
            return fut.result(timeout)
        except TimeoutError:
            raise RpcTimeout(f"ask {type(msg).__name__} -> {self.name}@{self.address} timed out after {timeout}s") from None

    def __reduce__(self):   # refs travel inside messages; they re-bind to the receiving process' env
        return (_rebind_ref, (self.address or (self._env.address if self._env else None), self.name))

    def __repr__(self) -> str:
        return f"RpcEndpointRef({self.name}@{self.address})"


_current_env = threading.local()
_default_env: Optional["RpcEnv"] = None


def _rebind_ref(address, name):
    env = getattr(_current_env, "env", None) or _default_env
    if env is not None and address is not None and env.address == RpcAddress(*address):
        return RpcEndpointRef(env, None, name)
    return RpcEndpointRef(env, RpcAddress(*address) if address is not None else None, name)


class _Inbox:
    def __init__(self, env: "RpcEnv", name: str, endpoint: RpcEndpoint):
        self.env, self.name, self.endpoint = env, name, endpoint
        self.q: "queue.Queue" = queue.Queue()
        self.thread = threading.Thread(target=self._loop, daemon=True, name=f"rpc-inbox-{env.name}-{name}")
        self.stopped = False

    def start(self) -> None:
        self.q.put(("start",))
        self.thread.start()

    def post(self, item: tuple) -> None:
        if not self.stopped:
            self.q.put(item)

    def stop(self) -> None:
        if not self.stopped:
            self.stopped = True
            self.q.put(("stop",))

    def _loop(self) -> None:
        _current_env.env = self.env
        ep = self.endpoint
        while True:
            item = self.q.get()
            kind = item[0]
            try:
                if kind == "start":
                    ep.on_start()
                elif kind == "stop":
                    ep.on_stop()
                    return
                elif kind == "oneway":
                    ep.receive(item[1], item[2])
                elif kind == "request":
                    ctx: RpcCallContext = item[2]
                    try:
                        ep.receive_and_reply(item[1], ctx)
                    except BaseException as exc:  # noqa: BLE001
                        ctx.send_failure(exc)
                elif kind == "connected":
                    ep.on_connected(item[1])
                elif kind == "disconnected":
                    ep.on_disconnected(item[1])
            except BaseException as exc:  # noqa: BLE001
                try:
                    ep.on_error(exc)
                except BaseException:  # noqa: BLE001
                    traceback.print_exc()


class _LocalCallContext(RpcCallContext):
    """ask() on an endpoint of the same env: no socket, the reply completes the future directly."""

    def __init__(self, fut: Future, sender: Optional[RpcAddress]):
        self._fut, self.sender_address, self._done = fut, sender, False

    def reply(self, value: Any) -> None:
        if not self._done:
            self._done = True
            self._fut.set_result(value)

    def send_failure(self, exc: BaseException) -> None:
        if not self._done:
            self._done = True
            self._fut.set_exception(exc if isinstance(exc, RpcError) else RpcRemoteFailure(repr(exc)))


class _Connection:
    """One socket + its reader thread.  ``peer`` is the remote env's advertised listening address (or the socket's
    peer name when the remote is a client-only env)."""

    def __init__(self, env: "RpcEnv", sock: socket.socket, peer: Optional[RpcAddress], dialed: bool):
        self.env, self.sock, self.peer, self.dialed = env, sock, peer, dialed
        self.send_lock = threading.Lock()
        self.pending: Dict[int, Future] = {}
        self.pending_lock = threading.Lock()
        self.closed = False
        self.reader: Optional[threading.Thread] = None
        self._enc = self._dec = None            # AES-GCM ciphers per direction once the handshake enabled encryption
        self._enc_ctr = self._dec_ctr = 0
        sock.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)

    def enable_encryption(self, secret: bytes, nonce: bytes, dialed: bool) -> None:
        """AES-256-GCM for every frame body after the handshake (``spark.network.crypto.enabled``, common/network-common
        crypto/): one key per direction from HKDF(secret, nonce); the GCM nonce is the frame counter, the frame type and
        request id are authenticated as associated data, so a tampered or replayed frame fails to decrypt."""
        from cryptography.hazmat.primitives import hashes
        from cryptography.hazmat.primitives.ciphers.aead import AESGCM
        from cryptography.hazmat.primitives.kdf.hkdf import HKDF
        okm = HKDF(algorithm=hashes.SHA256(), length=64, salt=nonce, info=b"afb-rpc-v1").derive(secret)
        c2s, s2c = AESGCM(okm[:32]), AESGCM(okm[32:])
        self._enc, self._dec = (c2s, s2c) if dialed else (s2c, c2s)

    def send_frame(self, ftype: int, req_id: int, body: bytes) -> None:
        if len(body) >= MAX_FRAME:
            raise RpcError(f"frame of {len(body)} bytes exceeds the 2 GiB frame limit")
        with self.send_lock:
            if self.closed:
                raise OSError("connection closed")
            if self._enc is not None:
                body = self._enc.encrypt(self._enc_ctr.to_bytes(12, "big"), bytes(body), struct.pack("!BQ", ftype, req_id))
                self._enc_ctr += 1
            hdr = _HDR.pack(len(body), ftype, req_id)
            if len(body) <= 65536:
                self.sock.sendall(hdr + bytes(body))
            else:                      # bulk path: no concatenation copy of a multi-MiB piece
                self.sock.sendall(hdr)
                self.sock.sendall(body)

    def recv_exact(self, n: int) -> bytearray:
        buf = bytearray(n)
        view = memoryview(buf)
        got = 0
        while got < n:
            r = self.sock.recv_into(view[got:], n - got)
            if r == 0:
                raise ConnectionError("peer closed")
            got += r
        return buf

    def recv_frame(self) -> Tuple[int, int, bytearray]:
        length, ftype, req_id = _HDR.unpack(self.recv_exact(_HDR.size))
        body = self.recv_exact(length) if length else bytearray()
        if self._dec is not None:
            from cryptography.exceptions import InvalidTag
            try:
                body = bytearray(self._dec.decrypt(self._dec_ctr.to_bytes(12, "big"), bytes(body), struct.pack("!BQ", ftype, req_id)))
            except InvalidTag:
                raise ConnectionError("frame failed authentication (tampered, replayed or wrong key)") from None
            self._dec_ctr += 1
        return ftype, req_id, body

    def start_reader(self) -> None:
        self.reader = threading.Thread(target=self._read_loop, daemon=True, name=f"rpc-conn-{self.env.name}-{self.peer}")
        self.reader.start()

    def _read_loop(self) -> None:
        _current_env.env = self.env
        try:
            while not self.closed:
                ftype, req_id, body = self.recv_frame()
                self.env._on_frame(self, ftype, req_id, body)
        except (ConnectionError, OSError, struct.error):
            pass
        except BaseException:  # noqa: BLE001  (a corrupt frame must not leave the connection half-alive)
            traceback.print_exc()
        finally:
            self.close()
            self.env._on_connection_closed(self)

    def close(self) -> None:
        if self.closed:
            return
        self.closed = True
        try:
            self.sock.shutdown(socket.SHUT_RDWR)
        except OSError:
            pass
        try:
            self.sock.close()
        except OSError:
            pass
        with self.pending_lock:
            pend, self.pending = self.pending, {}
        for fut in pend.values():
            if not fut.done():
                fut.set_exception(RpcDisconnected(f"connection to {self.peer} closed"))


class RpcEnv:
    """``RpcEnv.create(name, host, port)``.  ``port=0`` picks a free port; ``server=False`` makes a client-only env
    (it can still receive replies and server-initiated messages on the connections it dialed)."""

    _req_ids = itertools.count(1)

    def __init__(self, name: str = "rpc", host: str = "127.0.0.1", port: int = 0, server: bool = True,
                 secret: Optional[str] = None, connect_timeout: float = 10.0, encrypt: bool = False):
        global _default_env
        self.name = name
        if secret is None:
            secret = os.environ.get("AFB_RPC_SECRET") or None      # the secret never travels in argv (visible to `ps`)
        self.secret = secret.encode() if isinstance(secret, str) else secret
        if encrypt and not self.secret:
            raise ValueError("encrypt=True needs a shared secret (spark.authenticate.secret)")
        # with a secret, EVERY frame after the handshake is authenticated: AES-GCM per direction (a MAC-only mode would need a
        # second framing; encryption is cheap next to pickling), so `encrypt` is implied by `secret`
        self.encrypt = bool(encrypt) or bool(self.secret)
        self.connect_timeout = connect_timeout
        self._inboxes: Dict[str, _Inbox] = {}
        self._conns: Dict[RpcAddress, _Connection] = {}
        self._accepted: set = set()
        self._lock = threading.RLock()
        self._dial_locks: Dict[RpcAddress, threading.Lock] = {}
        self._stream_manager: Optional[Callable[[str], Optional[bytes]]] = None
        self._stopped = False
        self._server: Optional[socket.socket] = None
        self.address: Optional[RpcAddress] = None
        self.bytes_sent = 0
        self.bytes_received = 0
        if server:
            srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
            srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
            srv.bind((host, port))
            if not self.secret and host not in ("127.0.0.1", "localhost", "::1"):
                import warnings
                warnings.warn(f"RpcEnv {name!r} listens on {host} WITHOUT a shared secret: any peer that can reach the port can "
                              "send pickled messages (set spark.authenticate.secret / AFB_RPC_SECRET)", RuntimeWarning, stacklevel=2)
            srv.listen(128)
            self._server = srv
            self.address = RpcAddress(host, srv.getsockname()[1])
            threading.Thread(target=self._accept_loop, daemon=True, name=f"rpc-accept-{name}").start()
        if _default_env is None or _default_env._stopped:
            _default_env = self

    # ------------------------------------------------------------------ endpoints
    def setup_endpoint(self, name: str, endpoint: RpcEndpoint) -> RpcEndpointRef:
        with self._lock:
            if name in self._inboxes:
                raise RpcError(f"endpoint {name!r} already registered in env {self.name!r}")
            endpoint.rpc_env = self
            endpoint.self_ref = RpcEndpointRef(self, None, name)
            inbox = _Inbox(self, name, endpoint)
            self._inboxes[name] = inbox
        inbox.start()
        return endpoint.self_ref

    def stop_endpoint(self, name: str) -> None:
        with self._lock:
            inbox = self._inboxes.pop(name, None)
        if inbox is not None:
            inbox.stop()

    def endpoint_ref(self, address, name: str, verify: bool = False, timeout: float = 10.0) -> RpcEndpointRef:
        addr = RpcAddress.parse(address) if isinstance(address, str) else (RpcAddress(*address) if address is not None else None)
        if addr is not None and addr == self.address:
            addr = None
        ref = RpcEndpointRef(self, addr, name)
        if verify:
            if addr is None:
                if name not in self._inboxes:
                    raise RpcEndpointNotFound(name)
            elif not RpcEndpointRef(self, addr, "__verifier__").ask_sync(name, timeout):
                raise RpcEndpointNotFound(f"{name}@{addr}")
        return ref

    def set_stream_manager(self, fn: Callable[[str], Optional[bytes]]) -> None:
        """``fn(chunk_id) -> bytes-like | None`` serves ``fetch_chunk`` requests (OneForOneStreamManager's role)."""
        self._stream_manager = fn

    # ------------------------------------------------------------------ sending
    def _local_inbox(self, name: str) -> _Inbox:
        inbox = self._inboxes.get(name)
        if inbox is None:
            raise RpcEndpointNotFound(f"{name} (local env {self.name})")
        return inbox

    def _send_oneway(self, address: Optional[RpcAddress], name: str, msg: Any) -> None:
        if address is None:
            self._local_inbox(name).post(("oneway", msg, self.address))
            return
        body = cloudpickle.dumps((name, self.address, msg))
        self._connection(address).send_frame(F_ONEWAY, 0, body)
        self.bytes_sent += len(body)

    def _ask(self, address: Optional[RpcAddress], name: str, msg: Any) -> Future:
        fut: Future = Future()
        if address is None:
            try:
                self._local_inbox(name).post(("request", msg, _LocalCallContext(fut, self.address)))
            except RpcError as exc:
                fut.set_exception(exc)
            return fut
        try:
            conn = self._connection(address)
            req_id = next(RpcEnv._req_ids)
            with conn.pending_lock:
                conn.pending[req_id] = fut
            body = cloudpickle.dumps((name, self.address, msg))
            conn.send_frame(F_REQUEST, req_id, body)
            self.bytes_sent += len(body)
        except (OSError, RpcError) as exc:
            if not fut.done():
                fut.set_exception(exc if isinstance(exc, RpcError) else RpcDisconnected(f"{address}: {exc}"))
        return fut

    def fetch_chunk(self, address, chunk_id: str, timeout: float = 60.0) -> bytes:
        """Raw bulk fetch (``ChunkFetchRequest`` -> ``ChunkFetchSuccess``): the body is the block's bytes, unpickled."""
        addr = RpcAddress.parse(address) if isinstance(address, str) else RpcAddress(*address)
        if addr == self.address:
            data = self._stream_manager(chunk_id) if self._stream_manager else None
            if data is None:
                raise RpcError(f"chunk {chunk_id!r} not found")
            return bytes(data)
        fut: Future = Future()
        conn = self._connection(addr)
        req_id = next(RpcEnv._req_ids)
        with conn.pending_lock:
            conn.pending[req_id] = fut
        try:
            conn.send_frame(F_CHUNK_REQ, req_id, chunk_id.encode())
        except OSError as exc:
            raise RpcDisconnected(f"{addr}: {exc}") from None
        try:
            return fut.result(timeout)
        except TimeoutError:
            raise RpcTimeout(f"fetch of chunk {chunk_id!r} from {addr} timed out after {timeout}s") from None

    # ------------------------------------------------------------------ connections
    def _connection(self, address: RpcAddress) -> _Connection:
        with self._lock:
            if self._stopped:
                raise RpcError(f"env {self.name} is stopped")
            conn = self._conns.get(address)
            if conn is not None and not conn.closed:
                return conn
            dial_lock = self._dial_locks.setdefault(address, threading.Lock())
        with dial_lock:
            with self._lock:
                conn = self._conns.get(address)
                if conn is not None and not conn.closed:
                    return conn
            try:
                sock = socket.create_connection((address.host, address.port), timeout=self.connect_timeout)
            except OSError as exc:
                raise RpcDisconnected(f"cannot connect to {address}: {exc}") from None
            sock.settimeout(self.connect_timeout)
            conn = _Connection(self, sock, address, dialed=True)
            try:
                ftype, _, body = conn.recv_frame()                   # server speaks first: challenge (nonce)
                if ftype != F_CHALLENGE:
                    raise RpcError(f"bad handshake from {address}")
                nonce = bytes(body)
                # MUTUAL, direction-labelled, session-bound authentication: each side proves the secret over BOTH nonces with
                # its own label, so a proof cannot be relayed to another server (different server nonce) nor reflected back
                # (different label), and a rogue server that does not know the secret is rejected before any reply is unpickled
                cnonce = os.urandom(16)
                proof = hmac.new(self.secret, b"afb-rpc-client|" + nonce + cnonce, hashlib.sha256).digest() if self.secret else b""
                conn.send_frame(F_HELLO, 0, json.dumps({"addr": list(self.address) if self.address else None,
                                                        "proof": proof.hex(), "cnonce": cnonce.hex(), "enc": self.encrypt}).encode())
                ftype, _, body = conn.recv_frame()
                status, _, sproof = bytes(body).partition(b"|")
                if ftype != F_HELLO_ACK or status not in (b"ok", b"ok+enc"):
                    raise RpcError(f"{address} rejected the handshake: {bytes(body).decode(errors='replace')}")
                if self.secret:
                    want = hmac.new(self.secret, b"afb-rpc-server|" + nonce + cnonce, hashlib.sha256).hexdigest().encode()
                    if not hmac.compare_digest(want, sproof):
                        raise RpcError(f"{address} did not prove knowledge of the shared secret (rogue or misconfigured server)")
                body = status
                if bytes(body) == b"ok+enc":
                    conn.enable_encryption(self.secret, nonce, dialed=True)
                elif self.encrypt:
                    raise RpcError(f"{address} refused encryption")
            except (OSError, ConnectionError, struct.error) as exc:
                conn.close()
                raise RpcDisconnected(f"handshake with {address} failed: {exc}") from None
            except RpcError:
                conn.close()
                raise
            sock.settimeout(None)
            with self._lock:
                self._conns[address] = conn
            conn.start_reader()
            return conn

    def _accept_loop(self) -> None:
        srv = self._server
        while not self._stopped:
            try:
                sock, peer = srv.accept()
            except OSError:
                return
            threading.Thread(target=self._handshake_accepted, args=(sock, peer), daemon=True,
                             name=f"rpc-handshake-{self.name}").start()

    def _handshake_accepted(self, sock: socket.socket, peer) -> None:
        sock.settimeout(self.connect_timeout)
        conn = _Connection(self, sock, RpcAddress(peer[0], peer[1]), dialed=False)
        try:
            nonce = os.urandom(16)
            conn.send_frame(F_CHALLENGE, 0, nonce)
            ftype, _, body = conn.recv_frame()
            if ftype != F_HELLO:
                raise RpcError("expected HELLO")
            # the HELLO is parsed as JSON, never unpickled: nothing from an unauthenticated peer reaches the unpickler
            hello = json.loads(bytes(body).decode())
            advertised = tuple(hello["addr"]) if hello.get("addr") else None
            proof, wants_enc = bytes.fromhex(hello.get("proof", "")), bool(hello.get("enc", False))
            cnonce = bytes.fromhex(hello.get("cnonce", ""))
            if advertised is not None and not (isinstance(advertised[0], str) and isinstance(advertised[1], int)):
                raise RpcError("malformed HELLO")
            sproof = b""
            if self.secret is not None:
                want = hmac.new(self.secret, b"afb-rpc-client|" + nonce + cnonce, hashlib.sha256).digest()
                if len(cnonce) != 16 or not hmac.compare_digest(want, proof):
                    conn.send_frame(F_HELLO_ACK, 0, b"authentication failed")
                    raise RpcError("authentication failed")
                sproof = b"|" + hmac.new(self.secret, b"afb-rpc-server|" + nonce + cnonce, hashlib.sha256).hexdigest().encode()
            if self.encrypt and not wants_enc:
                conn.send_frame(F_HELLO_ACK, 0, b"this endpoint requires encryption")
                raise RpcError("peer did not ask for encryption")
            if (wants_enc or self.encrypt) and self.secret is not None:
                conn.send_frame(F_HELLO_ACK, 0, b"ok+enc" + sproof)
                conn.enable_encryption(self.secret, nonce, dialed=False)
            else:
                conn.send_frame(F_HELLO_ACK, 0, b"ok" + sproof)
        except BaseException:  # noqa: BLE001
            conn.close()
            return
        sock.settimeout(None)
        if advertised is not None:
            conn.peer = RpcAddress(*advertised)
        with self._lock:
            if self._stopped:
                conn.close()
                return
            self._accepted.add(conn)
            # a peer that advertises a listening address can be answered over the connection it opened
            if advertised is not None and (conn.peer not in self._conns or self._conns[conn.peer].closed):
                self._conns[conn.peer] = conn
            inboxes = list(self._inboxes.values())
        conn.start_reader()
        for ib in inboxes:
            ib.post(("connected", conn.peer))

    def _on_connection_closed(self, conn: _Connection) -> None:
        with self._lock:
            self._accepted.discard(conn)
            if self._conns.get(conn.peer) is conn:
                del self._conns[conn.peer]
            inboxes = list(self._inboxes.values())
            stopped = self._stopped
        if not stopped and conn.peer is not None:
            for ib in inboxes:
                ib.post(("disconnected", conn.peer))

    # ------------------------------------------------------------------ receiving
    def _on_frame(self, conn: _Connection, ftype: int, req_id: int, body: bytearray) -> None:
        self.bytes_received += len(body)
        if ftype in (F_RESPONSE, F_FAILURE, F_CHUNK_OK, F_CHUNK_FAIL):
            with conn.pending_lock:
                fut = conn.pending.pop(req_id, None)
            if fut is None or fut.done():
                return
            if ftype == F_RESPONSE:
                try:
                    fut.set_result(pickle.loads(body))
                except BaseException as exc:  # noqa: BLE001
                    fut.set_exception(RpcError(f"cannot decode reply: {exc!r}"))
            elif ftype == F_CHUNK_OK:
                fut.set_result(bytes(body))
            elif ftype == F_FAILURE:
                fut.set_exception(RpcRemoteFailure(pickle.loads(body)))
            else:
                fut.set_exception(RpcError(bytes(body).decode(errors="replace")))
            return
        if ftype == F_CHUNK_REQ:
            chunk_id = bytes(body).decode()
            try:
                data = self._stream_manager(chunk_id) if self._stream_manager else None
            except BaseException as exc:  # noqa: BLE001
                data, err = None, repr(exc)
            else:
                err = f"chunk {chunk_id!r} not found"
            try:
                if data is None:
                    conn.send_frame(F_CHUNK_FAIL, req_id, err.encode())
                else:
                    conn.send_frame(F_CHUNK_OK, req_id, data)
                    self.bytes_sent += len(data)
            except OSError:
                pass
            return
        if ftype in (F_ONEWAY, F_REQUEST):
            if self._stopped:          # shutting down: the peer will see the connection close, not a misleading "endpoint not found"
                return
            try:
                name, sender, msg = pickle.loads(body)
            except BaseException as exc:  # noqa: BLE001
                if ftype == F_REQUEST:
                    RpcCallContext(conn, req_id, conn.peer).send_failure(RpcError(f"cannot decode request: {exc!r}"))
                else:
                    traceback.print_exc()
                return
            sender = RpcAddress(*sender) if sender is not None else conn.peer
            if name == "__verifier__" and ftype == F_REQUEST:
                RpcCallContext(conn, req_id, sender).reply(msg in self._inboxes)
                return
            inbox = self._inboxes.get(name)
            if inbox is None:
                if ftype == F_REQUEST:
                    RpcCallContext(conn, req_id, sender).send_failure(RpcEndpointNotFound(f"{name}@{self.address}"))
                return
            if ftype == F_ONEWAY:
                inbox.post(("oneway", msg, sender))
            else:
                inbox.post(("request", msg, RpcCallContext(conn, req_id, sender)))

    # ------------------------------------------------------------------ lifecycle
    def shutdown(self) -> None:
        global _default_env
        with self._lock:
            if self._stopped:
                return
            self._stopped = True
            inboxes = list(self._inboxes.values())
            self._inboxes.clear()
            conns = set(self._conns.values()) | set(self._accepted)
            self._conns.clear()
            self._accepted.clear()
        for ib in inboxes:
            ib.stop()
        if self._server is not None:
            try:                      # wake the accept() thread: a plain close() leaves it blocked on the fd NUMBER, and if a
                self._server.shutdown(socket.SHUT_RDWR)   # new listener re-uses that number it would steal its connections
            except OSError:
                pass
            try:
                self._server.close()
            except OSError:
                pass
        for c in conns:
            c.close()
        if _default_env is self:
            _default_env = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.shutdown()
