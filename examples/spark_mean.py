#!/usr/bin/env python
"""The reference's toy program ``SparkMean`` (examples/src/main/scala/org/apache/spark/examples/SparkMean.scala:32-70): the mean
over rows of the per-row feature sums of a LIBSVM file, recomputed for a number of iterations.

    python examples/spark_mean.py <file> <num columns> <num rows> <num partitions> <num iterations>
    python examples/spark_mean.py            # no arguments: a small synthetic file is generated first

Same five positional arguments and the same progress lines; the record-level ``map(...).reduce(_+_)`` of the reference is also
shown in its columnar form (one vectorised sum per partition).
"""
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from asyncframework_b200.compat import MLUtils, SparkConf, SparkContext  # noqa: E402


def main(argv):
    print("input format: [file name] [num columns] [num rows] [num partitions] [num iterations]")
    if len(argv) < 5:
        from asyncframework_b200.core.scheduler import Scheduler
        from asyncframework_b200.data import LinearDataGenerator
        d = tempfile.mkdtemp()
        with Scheduler(2) as s:
            MLUtils.saveAsLibSVMFile(LinearDataGenerator.generate_linear_dataset(s, 500, 10, 0.1, 2, seed=3), os.path.join(d, "toy"))
        argv = [os.path.join(d, "toy"), "10", "500", "2", "3"]
    fname, d, N, num_part, iters = argv[0], int(argv[1]), int(argv[2]), int(argv[3]), int(argv[4])
    sc = SparkContext(SparkConf(False).setAppName("Average").setMaster("local[2]"))
    points = MLUtils.loadLibSVMFile(sc, fname).repartition(num_part)
    average = 0.0
    for i in range(1, iters + 1):
        print(f"******** iteration {i}********")
        average = points.map(lambda p: float(p.features.to_dense().values.sum())).reduce(lambda a, b: a + b) / N
    columnar = points.map_batch(lambda s: float(s.to_dense().X[:, :d].sum())).reduce(lambda a, b: a + b) / N
    print("The Final results is: ")
    print(average)
    assert abs(average - columnar) < 1e-9 * max(1.0, abs(average))
    sc.stop()
    return average


if __name__ == "__main__":
    main(sys.argv[1:])
