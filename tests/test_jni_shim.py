"""The JNI shim sources (spark-rapids-jni_b200/jni/) compile against the stub JNI / cudf headers and define exactly
the symbols of the reference's RowConversionJni.cpp:23-124, hash/HashJni.cpp:26-79 and KudoGpuSerializerJni.cpp:22-140, plus the
bindings of the two new plugin-side classes (HashPartition, UnsafeRowConversion)."""
import os
import shutil
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JNI = os.path.join(ROOT, "spark-rapids-jni_b200", "jni")

EXPECTED = {
    "RowConversionJni.cpp": ["Java_com_nvidia_spark_rapids_jni_RowConversion_convertToRows",
                             "Java_com_nvidia_spark_rapids_jni_RowConversion_convertToRowsFixedWidthOptimized",
                             "Java_com_nvidia_spark_rapids_jni_RowConversion_convertFromRows",
                             "Java_com_nvidia_spark_rapids_jni_RowConversion_convertFromRowsFixedWidthOptimized"],
    "HashJni.cpp": ["Java_com_nvidia_spark_rapids_jni_Hash_getMaxStackDepth", "Java_com_nvidia_spark_rapids_jni_Hash_murmurHash32",
                    "Java_com_nvidia_spark_rapids_jni_Hash_xxhash64", "Java_com_nvidia_spark_rapids_jni_Hash_hiveHash"],
    # kudo/KudoGpuSerializer.java's two natives (KudoGpuSerializerJni.cpp:22-140), flat schemas
    "KudoGpuSerializerJni.cpp": ["Java_com_nvidia_spark_rapids_jni_kudo_KudoGpuSerializer_splitAndSerializeToDevice",
                                 "Java_com_nvidia_spark_rapids_jni_kudo_KudoGpuSerializer_assembleFromDeviceRawNative"],
    # new classes for the plugin-side steps (no reference counterpart): INTEGRATION.md 2c
    "HashPartitionJni.cpp": ["Java_com_nvidia_spark_rapids_jni_HashPartition_hashPartition"],
    "UnsafeRowConversionJni.cpp": ["Java_com_nvidia_spark_rapids_jni_UnsafeRowConversion_convertToRows",
                                   "Java_com_nvidia_spark_rapids_jni_UnsafeRowConversion_convertFromRows"],
}


@pytest.mark.parametrize("src", sorted(EXPECTED))
def test_shim_compiles_and_defines_the_reference_symbols(src):
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("g++ not available")
    with tempfile.TemporaryDirectory() as td:
        obj = os.path.join(td, src.replace(".cpp", ".o"))
        r = subprocess.run([gxx, "-std=c++17", "-Wall", "-Werror", "-fPIC", "-DSRJ_JNI_STUBS", "-c", os.path.join(JNI, src), "-o", obj],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        syms = subprocess.run(["nm", "-g", "--defined-only", obj], capture_output=True, text=True).stdout
        defined = {l.split()[-1] for l in syms.splitlines() if " T " in l}
        assert set(EXPECTED[src]) <= defined, set(EXPECTED[src]) - defined
        assert not [s for s in defined if s.startswith("Java_") and s not in EXPECTED[src]]
