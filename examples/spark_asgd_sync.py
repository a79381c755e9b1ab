#!/usr/bin/env python
"""SparkASGDSync transcribed line by line with the reference's own names (asyncframework_b200.compat).
Compare with ASYNCsamples/src/main/scala/org/apache/async/SparkASGDSync.scala:150-272.

    python examples/spark_asgd_sync.py [libsvm_path] [d]
"""
import math
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # run from a source checkout

import torch

from asyncframework_b200.compat import ASYNCcontext, MLUtils, SparkContext, Vectors, axpyOp, dotOp
from asyncframework_b200.data import LinearDataGenerator, save_as_libsvm_file


def gradfun(p, w):                                       # SparkASGDThread.scala:420-435
    grad = Vectors.zeros(w.size)
    diff = dotOp(p.features, w) - p.label
    axpyOp(diff, p.features, grad)
    return grad


def main(path=None, d=10, numPart=4, numIteration=30, gamma=0.5, b=0.3, InputSeed=42):
    sc = SparkContext("local[4]")
    sc.setLogLevel("ERROR")
    if path is None:                                     # no file given: write a small synthetic LIBSVM directory
        path = tempfile.mkdtemp() + "/data"
        save_as_libsvm_file(LinearDataGenerator.generate_linear_dataset(sc.scheduler, 800, d, 0.1, 2, seed=1), path)
    points = MLUtils.loadLibSVMFile(sc, path, d).repartition(numPart)
    N = points.count()
    w = Vectors.zeros(d)
    AC = ASYNCcontext(numPart, sc)
    pointsIndexed = points.zipWithIndex().cache()
    for k in range(numIteration):
        a = sc.broadcast(w)
        IndexGrad = pointsIndexed.sample(False, b, InputSeed + k + 1).map(lambda x, a=a: gradfun(x[0], a.value))
        IndexGrad.ASYNCreduce(lambda x, y: axpyOp(1.0, x, y), AC)
        accGrad = Vectors.zeros(d)
        got = 0
        while got < numPart:                              # wait for numPart results (SparkASGDSync.scala:234-266)
            info = AC.ASYNCcollectAll(timeout=5.0)
            if info is None:
                break
            axpyOp(1.0, info.gettaskResult(), accGrad)
            got += 1
        axpyOp(-gamma / math.sqrt(k + 1) / (b * N), accGrad, w)
    sc.set_mode(0)
    obj = pointsIndexed.map(lambda x: (dotOp(x[0].features, w) - x[0].label) ** 2).reduce(lambda p, q: p + q) / N
    print("objective", obj)
    sc.stop()
    return obj


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else None, int(sys.argv[2]) if len(sys.argv) > 2 else 10)
