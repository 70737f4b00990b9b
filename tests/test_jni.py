"""The JNI surface (reference encoder_jni.c:5-22, `make jni`): Java_com_example_Encoder_mpegEncodeProcedure
compiled from OUR encoder_jni.c, unchanged, against a stand-in <jni.h> (tests/host/jni.h: this image has no
JDK) and called through a fake JNIEnv (tests/host/jni_harness.c).  Checks: same result as calling
mpeg_encode_procedure directly, every Java string pinned once and released once with the pointer it was
given."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(HERE, "_build", "libm1jni_test.so")


@pytest.fixture(scope="module")
def jni():
    from ec504_imageencoder_b200 import _native
    _native.build_cuda()
    subprocess.run(["make", "-s", "-C", ROOT, "sharedlib"], check=True)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    pkg = os.path.join(ROOT, "ec504_imageencoder_b200")
    subprocess.run(["gcc", "-O2", "-g", "-fPIC", "-shared", "-Wall", "-I", os.path.join(HERE, "host"), "-I", ROOT,
                    "-o", OUT, os.path.join(ROOT, "encoder_jni.c"), os.path.join(HERE, "host", "jni_harness.c"),
                    "-L", pkg, "-lencoder", "-lm1cu", f"-Wl,-rpath,{pkg}", "-lm"], check=True, cwd=ROOT)
    lib = C.CDLL(OUT)
    lib.m1jni_call.restype = C.c_int
    lib.m1jni_call.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_int)]
    return lib


def _call(lib, images, streams, video, q=12):
    counts = (C.c_int * 8)()
    rc = lib.m1jni_call(str(images).encode(), str(streams).encode(), str(video).encode(), q, counts)
    return rc, list(counts)


def test_jni_exports_the_reference_symbol(jni):
    assert hasattr(jni, "Java_com_example_Encoder_mpegEncodeProcedure")


def test_jni_return_codes_and_string_lifetime_without_gpu(jni, tmp_path):
    """The paths of the driver that need no GPU, through the JNI entry point."""
    # unwritable video path -> 1
    rc, c = _call(jni, tmp_path / "imgs", tmp_path / "bs", tmp_path / "missing_dir" / "v.mpeg")
    assert rc == 1 and c == [1, 1, 1, 1, 1, 1, 0, 0]
    # missing images folder -> created, 0, pack + system header written
    rc, c = _call(jni, tmp_path / "imgs", tmp_path / "bs", tmp_path / "v.mpeg")
    assert rc == 0 and c == [1, 1, 1, 1, 1, 1, 0, 0]
    assert (tmp_path / "imgs").is_dir() and (tmp_path / "bs").is_dir() and (tmp_path / "v.mpeg").stat().st_size == 27
    # empty folder -> -1
    rc, c = _call(jni, tmp_path / "imgs", tmp_path / "bs", tmp_path / "v.mpeg")
    assert rc == -1 and c == [1, 1, 1, 1, 1, 1, 0, 0]


@pytest.mark.gpu
def test_jni_call_equals_direct_call(jni, tmp_path):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    PIL = pytest.importorskip("PIL.Image")
    from ec504_imageencoder_b200 import hostlib
    z = np.load(os.path.join(HERE, "golden", "refcompat_inputs.npz"))
    imgs = tmp_path / "images"
    imgs.mkdir()
    pics = z["images"][:6]
    for i, im in enumerate(pics):                                    # the reference fixture's pictures (400 x 144 crop)
        full = np.zeros((600, 400, 3), np.uint8)
        full[:144] = im
        PIL.fromarray(full).save(str(imgs / f"f{i}.jpg"), quality=95)
    if hostlib.mpeg_encode_procedure(str(imgs), str(tmp_path / "d"), str(tmp_path / "d.mpeg"), 12) != 0:
        pytest.skip("libencoder.so was built without stb_image.h")
    rc, c = _call(jni, imgs, tmp_path / "j", tmp_path / "j.mpeg", 12)
    assert rc == 0 and c == [1, 1, 1, 1, 1, 1, 0, 0]
    assert (tmp_path / "j.mpeg").read_bytes() == (tmp_path / "d.mpeg").read_bytes()
    assert len((tmp_path / "j.mpeg").read_bytes()) > 27 + len(pics) * 48
    for i in range(1, len(pics) + 1):
        assert (tmp_path / "j" / f"image_{i}.bit").read_bytes() == (tmp_path / "d" / f"image_{i}.bit").read_bytes()
