"""RPC environment: endpoints, send/ask, chunk fetch, disconnect detection, authentication
(the reference's RpcEnvSuite / NettyRpcEnvSuite scenarios, core/src/test/scala/org/apache/spark/rpc/RpcEnvSuite.scala)."""
import threading
import time

import pytest

from asyncframework_b200.rpc import (RpcDisconnected, RpcEndpoint, RpcEndpointNotFound, RpcEnv, RpcError,
                                     RpcRemoteFailure, RpcTimeout)


class Echo(RpcEndpoint):
    def __init__(self):
        self.got = []
        self.events = []
        self.started = threading.Event()

    def on_start(self):
        self.started.set()

    def receive(self, msg, sender):
        self.got.append((msg, sender))

    def receive_and_reply(self, msg, ctx):
        if msg == "boom":
            raise ValueError("kaboom")
        if msg == "never":
            self.got.append(("never", ctx.sender_address))
            return
        if isinstance(msg, tuple) and msg[0] == "callback":      # reply later, from another thread
            threading.Timer(0.05, lambda: ctx.reply(("late", msg[1]))).start()
            return
        ctx.reply(("echo", msg))

    def on_connected(self, address):
        self.events.append(("connected", address))

    def on_disconnected(self, address):
        self.events.append(("disconnected", address))


def _wait(pred, timeout=20.0):
    t0 = time.time()
    while time.time() - t0 < timeout:
        if pred():
            return True
        time.sleep(0.01)
    return False


def test_send_ask_local_and_remote():
    with RpcEnv("server") as server, RpcEnv("client") as client:
        ep = Echo()
        local = server.setup_endpoint("echo", ep)
        assert ep.started.wait(2)
        assert local.ask_sync("x") == ("echo", "x")                       # same-env short circuit
        ref = client.endpoint_ref(server.address, "echo", verify=True)
        assert ref.ask_sync({"a": [1, 2, 3]}) == ("echo", {"a": [1, 2, 3]})
        assert ref.ask_sync(("callback", 7)) == ("late", 7)
        ref.send("hello")
        assert _wait(lambda: any(m == "hello" for m, _ in ep.got))
        sender = [s for m, s in ep.got if m == "hello"][0]
        assert sender == client.address                                   # the callee can call back
        with pytest.raises(RpcRemoteFailure, match="kaboom"):
            ref.ask_sync("boom")
        with pytest.raises(RpcTimeout):
            ref.ask_sync("never", timeout=0.1)
        with pytest.raises(RpcEndpointNotFound):
            client.endpoint_ref(server.address, "nope", verify=True)
        with pytest.raises(RpcRemoteFailure, match="nope"):
            client.endpoint_ref(server.address, "nope").ask_sync(1)


def test_messages_to_one_endpoint_are_sequential_and_ordered():
    class Counter(RpcEndpoint):
        def __init__(self):
            self.seen, self.inside, self.overlap = [], 0, False

        def receive(self, msg, sender):
            self.inside += 1
            self.overlap |= self.inside > 1
            time.sleep(0.0005)
            self.seen.append(msg)
            self.inside -= 1

    with RpcEnv("s") as server, RpcEnv("c", server=False) as client:
        ep = Counter()
        server.setup_endpoint("ctr", ep)
        ref = client.endpoint_ref(server.address, "ctr")
        for i in range(200):
            ref.send(i)
        assert _wait(lambda: len(ep.seen) == 200)
        assert ep.seen == list(range(200)) and not ep.overlap


def test_server_can_call_back_a_client_only_env_and_refs_travel_in_messages():
    class Registry(RpcEndpoint):
        def __init__(self):
            self.refs = []

        def receive_and_reply(self, msg, ctx):
            self.refs.append(msg)           # an RpcEndpointRef pickled by the client
            ctx.reply(True)

    with RpcEnv("s") as server, RpcEnv("c") as client:
        reg = Registry()
        server.setup_endpoint("reg", reg)
        mine = Echo()
        my_ref = client.setup_endpoint("mine", mine)
        assert client.endpoint_ref(server.address, "reg").ask_sync(my_ref)
        back = reg.refs[0]
        assert back.address == client.address and back.name == "mine"
        back.send("from-server")
        assert _wait(lambda: any(m == "from-server" for m, _ in mine.got))
        assert back.ask_sync("q") == ("echo", "q")


def test_disconnect_fails_outstanding_asks_and_notifies_endpoints():
    server, client = RpcEnv("s"), RpcEnv("c")
    try:
        ep = Echo()
        server.setup_endpoint("echo", ep)
        ref = client.endpoint_ref(server.address, "echo")
        assert ref.ask_sync(1) == ("echo", 1)
        assert _wait(lambda: ("connected", client.address) in ep.events), f"no connected event: {ep.events}"
        cep = Echo()
        client.setup_endpoint("watch", cep)
        assert cep.started.wait(5)
        fut = ref.ask("never")
        assert _wait(lambda: any(m == "never" for m, _ in ep.got))       # the request is IN the endpoint when the server goes away
        server.shutdown()
        try:
            fut.result(20)
            raise AssertionError("the outstanding ask completed instead of failing")
        except RpcDisconnected:
            pass
        except BaseException as exc:  # noqa: BLE001
            raise AssertionError(f"outstanding ask failed with {type(exc).__name__}: {exc!r} instead of RpcDisconnected")
        assert _wait(lambda: any(e[0] == "disconnected" for e in cep.events)), f"no disconnected event on the client: {cep.events}"
        try:
            got = ref.ask_sync(2, timeout=15)
            raise AssertionError(f"ask to a dead server returned {got!r}")
        except RpcError:
            pass
    finally:
        client.shutdown()
        server.shutdown()


def test_server_sees_client_loss():
    with RpcEnv("s") as server:
        ep = Echo()
        server.setup_endpoint("echo", ep)
        client = RpcEnv("c")
        client.endpoint_ref(server.address, "echo").ask_sync("hi")
        addr = client.address
        client.shutdown()
        assert _wait(lambda: ("disconnected", addr) in ep.events)


def test_chunk_fetch_is_raw_and_handles_large_blocks():
    blocks = {"broadcast_1_piece0": bytes(range(256)) * 4096 * 5, "empty": b""}      # 5 MiB
    with RpcEnv("s") as server, RpcEnv("c", server=False) as client:
        server.set_stream_manager(blocks.get)
        assert client.fetch_chunk(server.address, "broadcast_1_piece0") == blocks["broadcast_1_piece0"]
        assert client.fetch_chunk(server.address, "empty") == b""
        with pytest.raises(RpcError, match="not found"):
            client.fetch_chunk(server.address, "missing")
        assert server.fetch_chunk(server.address, "empty") == b""          # local short circuit


def test_shared_secret_authentication():
    with RpcEnv("s", secret="s3cret") as server:
        server.setup_endpoint("echo", Echo())
        good = RpcEnv("good", server=False, secret="s3cret")
        assert good.endpoint_ref(server.address, "echo").ask_sync(1) == ("echo", 1)
        good.shutdown()
        bad = RpcEnv("bad", server=False, secret="wrong")
        with pytest.raises(RpcError, match="authentication failed"):
            bad.endpoint_ref(server.address, "echo").ask_sync(1, timeout=2)
        bad.shutdown()
        anon = RpcEnv("anon", server=False)
        with pytest.raises(RpcError):
            anon.endpoint_ref(server.address, "echo").ask_sync(1, timeout=2)
        anon.shutdown()


def test_concurrent_asks_from_many_threads():
    with RpcEnv("s") as server, RpcEnv("c", server=False) as client:
        server.setup_endpoint("echo", Echo())
        ref = client.endpoint_ref(server.address, "echo")
        out, errs = {}, []

        def work(t):
            try:
                for i in range(100):
                    assert ref.ask_sync((t, i)) == ("echo", (t, i))
                out[t] = True
            except BaseException as exc:  # noqa: BLE001
                errs.append(exc)
        ths = [threading.Thread(target=work, args=(t,)) for t in range(8)]
        [t.start() for t in ths]
        [t.join() for t in ths]
        assert not errs and len(out) == 8


def test_aes_gcm_encryption_hides_payloads_and_rejects_tampering():
    """spark.network.crypto.enabled: frames are encrypted + authenticated after the handshake."""
    import socket
    import threading
    marker = b"TOP-SECRET-PAYLOAD-0123456789"
    seen = bytearray()
    with RpcEnv("s", secret="k", encrypt=True) as server:
        server.setup_endpoint("echo", Echo())
        server.set_stream_manager({"blk": marker * 1000}.get)
        # a byte-copying proxy in the middle records everything on the wire
        lsock = socket.socket()
        lsock.bind(("127.0.0.1", 0))
        lsock.listen(1)
        flip = {"on": False}

        def pump(a, b, record):
            try:
                while True:
                    d = a.recv(65536)
                    if not d:
                        break
                    if record:
                        seen.extend(d)
                    if flip["on"] and len(d) > 40:
                        d = d[:30] + bytes([d[30] ^ 0x01]) + d[31:]
                        flip["on"] = False
                    b.sendall(d)
            except OSError:
                pass
            finally:
                for x in (a, b):
                    try:
                        x.close()
                    except OSError:
                        pass

        def proxy():
            c, _ = lsock.accept()
            u = socket.create_connection(tuple(server.address))
            threading.Thread(target=pump, args=(c, u, True), daemon=True).start()
            threading.Thread(target=pump, args=(u, c, True), daemon=True).start()
        threading.Thread(target=proxy, daemon=True).start()
        client = RpcEnv("c", server=False, secret="k", encrypt=True)
        via = ("127.0.0.1", lsock.getsockname()[1])
        ref = client.endpoint_ref(via, "echo")
        assert ref.ask_sync(marker) == ("echo", marker)
        assert client.fetch_chunk(via, "blk") == marker * 1000
        assert marker not in bytes(seen) and len(seen) > len(marker) * 1000          # ciphertext only on the wire
        flip["on"] = True                                                              # now corrupt one byte in flight
        with pytest.raises(RpcError):
            ref.ask_sync(b"x" * 100, timeout=3)                                        # authentication fails -> connection dropped
        client.shutdown()
        # a shared secret implies encryption: an env built with only ``secret=`` never sends plaintext frames
        implied = RpcEnv("p", server=False, secret="k")
        assert implied.encrypt
        assert implied.endpoint_ref(server.address, "echo").ask_sync(1, timeout=3) == ("echo", 1)
        implied.shutdown()
    with pytest.raises(ValueError):
        RpcEnv("x", server=False, encrypt=True)


def test_unauthenticated_peer_cannot_reach_the_unpickler():
    """The handshake is JSON; a pickle bomb sent as HELLO (or as the first frame) is rejected before any unpickling."""
    import pickle
    import socket
    import struct

    class Bomb:
        def __reduce__(self):
            return (os.system, ("touch /tmp/afb_pwned_marker",))
    import os
    try:
        os.remove("/tmp/afb_pwned_marker")
    except FileNotFoundError:
        pass
    with RpcEnv("s", secret="k") as server:
        server.setup_endpoint("echo", Echo())
        for ftype in (0, 3):                                # F_HELLO, F_REQUEST
            sock = socket.create_connection(tuple(server.address), timeout=3)
            sock.recv(64)                                    # the server's challenge
            body = pickle.dumps(Bomb())
            sock.sendall(struct.pack("!IBQ", len(body), ftype, 1) + body)
            sock.settimeout(3)
            try:
                while sock.recv(4096):
                    pass
            except OSError:
                pass
            sock.close()
        assert not os.path.exists("/tmp/afb_pwned_marker")
        good = RpcEnv("g", server=False, secret="k")
        assert good.endpoint_ref(server.address, "echo").ask_sync(5) == ("echo", 5)      # the server is still healthy
        good.shutdown()
