from .batch import Shard, concat_shards, padded_ld
from .dataset import HashPartitioner, RangePartitioner, ShardedDataset
from .libsvm import MLUtils, load_libsvm_file, parse_libsvm_record, save_as_libsvm_file
from .generators import (LinearDataGenerator, LogisticRegressionDataGenerator, generate_gd_input,
                         synthetic_dataset, synthetic_shard, SHAPES)

__all__ = ["Shard", "concat_shards", "padded_ld", "ShardedDataset", "HashPartitioner", "RangePartitioner", "MLUtils", "load_libsvm_file",
           "parse_libsvm_record", "save_as_libsvm_file", "LinearDataGenerator",
           "LogisticRegressionDataGenerator", "generate_gd_input", "synthetic_dataset", "synthetic_shard", "SHAPES"]
