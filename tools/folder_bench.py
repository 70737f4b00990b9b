#!/usr/bin/env python3
"""Wall time of the drop-in entry point mpeg_encode_procedure() on a folder of JPEGs:
N synthetic 1920x1080 pictures written as JPEG files (cv2), then the C driver in three settings:
  decode-first     M1_DECODE_THREADS=0   the reference's order of work: decode every file, then encode (round 1)
  pipeline 1       M1_DECODE_THREADS=1   one decode thread overlapped with upload / encode / file output
  pipeline (auto)  default               min(16, cores) decode threads
M1_MODE=full (whole pictures), .bit side files on and off.  Prints one JSON object."""
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cv2  # noqa: E402
import numpy as np  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    import oracle
    from ec504_imageencoder_b200 import hostlib
    port = oracle.Port()
    work = tempfile.mkdtemp(prefix="m1_folder_")
    imgs = os.path.join(work, "images")
    os.makedirs(imgs)
    base = [port.synth_rgb(12345, f, 1920, 1080, 0) for f in range(8)]
    for i in range(n):                                           # 8 distinct pictures, written n / 8 times each
        cv2.imwrite(os.path.join(imgs, f"p{i:05d}.jpg"), base[i % 8][..., ::-1], [cv2.IMWRITE_JPEG_QUALITY, 90])
    jpeg_bytes = sum(os.path.getsize(os.path.join(imgs, f)) for f in os.listdir(imgs))
    os.environ["M1_MODE"] = "full"
    out = {"pictures": n, "width": 1920, "height": 1080, "jpeg_megabytes": jpeg_bytes / 1e6, "host_cores": len(os.sched_getaffinity(0)),
           "runs": []}
    for bit in ("1", "0"):
        for label, thr in (("decode-first (M1_DECODE_THREADS=0)", "0"), ("pipeline, 1 decode thread", "1"), ("pipeline, auto threads", None)):
            os.environ["M1_BIT_FILES"] = bit
            if thr is None:
                os.environ.pop("M1_DECODE_THREADS", None)
            else:
                os.environ["M1_DECODE_THREADS"] = thr
            dst = os.path.join(work, "out")
            shutil.rmtree(dst, ignore_errors=True)
            os.makedirs(dst)
            devnull = os.open(os.devnull, os.O_WRONLY)
            saved = os.dup(1)
            os.dup2(devnull, 1)                                  # the driver prints one line per picture
            t0 = time.perf_counter()
            rc = hostlib.mpeg_encode_procedure(imgs, dst, os.path.join(dst, "v.mpeg"), 12)
            dt = time.perf_counter() - t0
            os.dup2(saved, 1)
            os.close(devnull); os.close(saved)
            size = os.path.getsize(os.path.join(dst, "v.mpeg"))
            out["runs"].append({"setting": label, "bit_files": bit == "1", "rc": rc, "seconds": dt, "pictures_per_s": n / dt,
                                "video_bytes": size})
    shutil.rmtree(work, ignore_errors=True)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
