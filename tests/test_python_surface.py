"""CPU check of the host mirror's surface for the 8(f) paths: the modules import without a GPU, carry the names of the
classes they mirror (kudo/KudoGpuSerializer.java, cudf Table.partition / PartitionedTable, RowConversion-shaped
UnsafeRowConversion) and bind only symbols that include/srj_b200.h declares."""
import inspect
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spark-rapids-jni_b200"))


def test_modules_and_names():
    from srj_b200 import kudo, partitioning, shuffle, unsaferow
    assert {"splitAndSerializeToDevice", "assembleFromDeviceRaw"} <= set(dir(kudo.KudoGpuSerializer))       # KudoGpuSerializer.java
    assert {"getTable", "getPartitions"} <= set(dir(partitioning.PartitionedTable))                          # ai.rapids.cudf.PartitionedTable
    assert {"partition", "partitionIds"} <= set(dir(partitioning.HashPartitioner)) and callable(partitioning.partition)
    assert {"convertToRows", "convertFromRows", "layout"} <= set(dir(unsaferow.UnsafeRowConversion))          # RowConversion.java:120-174 shape
    assert list(inspect.signature(shuffle.exchange_partitions).parameters)[:3] == ["buf", "part_offsets", "parts_per_rank"]
    assert "shuffle" in dir(shuffle.ShuffleExchange)


def test_every_bound_symbol_is_declared_in_the_header():
    from srj_b200 import _native as N
    header = open(os.path.join(ROOT, "include", "srj_b200.h")).read()
    declared = set(re.findall(r"SRJ_API\s+[\w\s\*]+?\b(srj_\w+)\s*\(", header))
    used = set()
    for mod in ("kudo.py", "partitioning.py", "unsaferow.py", "hostpath.py", "sharding.py", "__init__.py"):
        src = open(os.path.join(ROOT, "spark-rapids-jni_b200", "srj_b200", mod)).read()
        used |= set(re.findall(r"\blib(?:\(\))?\.(srj_\w+)", src))
    assert used and used <= declared, used - declared
    assert declared == set(N.SIGNATURES) if hasattr(N, "SIGNATURES") else True
