"""AsyncConf (SparkConfSuite scenarios) and the submit CLI (SparkSubmitSuite argument handling)."""
import os
import subprocess
import sys

import pytest

from asyncframework_b200 import submit
from asyncframework_b200.conf import AsyncConf, ENTRIES, parse_size, parse_time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_size_and_time_strings():
    assert parse_size("512m") == 512 << 20 and parse_size("2g") == 2 << 30 and parse_size("4096") == 4096
    assert parse_size("15", "g") == 15 << 30 and parse_size("1.5k") == 1536
    assert parse_time("120s") == 120.0 and parse_time("100ms") == 0.1 and parse_time("2min") == 120.0 and parse_time("5") == 5.0
    for bad in ("12x", "", "abc"):
        with pytest.raises(ValueError):
            parse_size(bad)
        with pytest.raises(ValueError):
            parse_time(bad)


def test_typed_entries_defaults_validation_and_deprecation(tmp_path):
    c = AsyncConf(load_defaults=False)
    assert c.get("spark.task.maxFailures") == 4 and c.get("spark.broadcast.blockSize") == 4 << 20
    c.set("spark.executor.memory", "13g").set("spark.network.timeout", "90s").set("spark.speculation", "true")
    assert c.get("spark.executor.memory") == 13 << 30 and c.get("spark.network.timeout") == 90.0 and c.get("spark.speculation") is True
    with pytest.raises(ValueError, match="spark.task.maxFailures"):
        c.set("spark.task.maxFailures", "0")
    with pytest.raises(ValueError, match="spark.speculation"):
        c.set("spark.speculation", "maybe")
    c.set("spark.akka.timeout", "30s")                                     # deprecated spelling is translated
    assert c.get("spark.network.timeout") == 30.0 and "deprecated" in c.warnings[0]
    c.set("my.custom.key", "v")                                            # unknown keys are kept verbatim
    assert c.get("my.custom.key") == "v" and ("my.custom.key", "v") in c.get_all()
    with pytest.raises(KeyError):
        c.get("not.set")
    c.set("spark.authenticate", "true")
    with pytest.raises(ValueError, match="secret"):
        c.validate()
    c.set("spark.authenticate.secret", "s3")
    assert "redacted" in c.to_debug_string() and "s3" not in c.to_debug_string()
    lim = c.validate().runtime_limits()
    assert lim.auth_secret == "s3" and lim.storage_memory == int((13 << 30) * 0.6) and lim.compression_codec == "lz4"
    f = tmp_path / "d.conf"
    f.write_text("# comment\nspark.executor.cores 2\nspark.ui.port=0\nspark.io.compression.codec  org.apache.spark.io.ZStdCompressionCodec\n")
    c2 = AsyncConf(load_defaults=False).set("spark.executor.cores", 8).load_properties_file(str(f))
    assert c2.get("spark.executor.cores") == 8 and c2.get("spark.ui.port") == 0           # explicit set wins over the file
    assert c2.get("spark.io.compression.codec") == "zstd"
    assert len(AsyncConf.describe().splitlines()) == len(ENTRIES) + 2


def test_submit_precedence_and_reference_class_names(tmp_path, monkeypatch):
    props = tmp_path / "p.conf"
    props.write_text("spark.executor.cores 4\nspark.app.name from-file\nspark.task.maxFailures 7\n")
    ns = submit.build_parser().parse_args(["--master", "local[2]", "--conf", "spark.task.maxFailures=2", "--properties-file", str(props),
                                           "--executor-memory", "15g", "--driver-memory", "20g", "--total-executor-cores", "16",
                                           "--class", "org.apache.async.SparkASGDThread", "app.jar", "a", "b"])
    conf = submit.resolve_conf(ns)
    assert conf.get("spark.master") == "local[2]" and conf.get("spark.task.maxFailures") == 2          # flag > --conf > file
    assert conf.get("spark.app.name") == "from-file" and conf.get("spark.executor.cores") == 4
    assert conf.get("spark.executor.memory") == 15 << 30 and conf.get("spark.executor.instances") == 4  # 16 cores / 4 per executor
    env = submit.export_conf(conf, {})
    assert env["AFB_CONF_spark_task_maxFailures"] == "2"
    monkeypatch.setenv("AFB_CONF_spark_task_maxFailures", "2")
    assert AsyncConf().get("spark.task.maxFailures") == 2                     # what a submitted script sees
    assert submit.CLASS_TO_DRIVER["org.apache.async.SparkASAGASync"] == "saga_sync"
    with pytest.raises(SystemExit):
        submit.main(["--class", "no.such.Class", "x"])


def test_submit_runs_a_reference_style_command_line_end_to_end(tmp_path):
    """README.md:46 shaped command: --class + memory flags + jar placeholder + the 11 positional sync arguments."""
    from asyncframework_b200.data.generators import LinearDataGenerator
    from asyncframework_b200.data.libsvm import save_as_libsvm_file
    from asyncframework_b200.core.scheduler import Scheduler
    with Scheduler(2) as s:
        ds = LinearDataGenerator.generate_linear_dataset(s, 200, 6, 0.1, 2, seed=1)
        save_as_libsvm_file(ds, str(tmp_path / "lin.libsvm"))
    cmd = [sys.executable, "-m", "asyncframework_b200.submit", "--class", "org.apache.async.SparkASGDSync", "--master", "local-cluster[2]",
           "--executor-memory", "1g", "--conf", "spark.task.maxFailures=2", "async.jar",
           str(tmp_path) + "/", "lin.libsvm", "6", "200", "2", "8", "0.5", "0.5", "4", "0", "42", "--backend=cpu", "--quiet=true"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240, cwd=ROOT, env={**os.environ, "PYTHONPATH": ROOT})
    assert out.returncode == 0, out.stderr[-2000:]
    assert "Spark SGD application started" in out.stdout and "num iterations: 8" in out.stdout and "Elapsed time(ms): " in out.stdout
