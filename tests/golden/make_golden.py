"""Copies the reference's shipped lossless example bitstreams into tests/golden/.

Run in the build container (where /root/reference exists):
    python tests/golden/make_golden.py
The .flo files are the reference encoder's own output (reflo 0.1.2, level 5,
checked in under Examples/); they are the only bit-level pins the reference
holds for the lossless encode path.  audio.wav is
1 s of all-zero IEEE-float stereo, stored gzipped (it is 352 KB of zeros).
"""
import gzip
import hashlib
import json
import os
import shutil
import sys

REF = os.environ.get("FLO_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
LOSSLESS = [
    "audio_lossless.flo", "chord_cmajor_stereo.flo", "click_track_120bpm.flo", "dtmf_tones.flo",
    "hires_96khz.flo", "multitone_stereo.flo", "silence_1sec.flo", "sine_440hz_mono.flo",
    "sweep_20_20k.flo", "telephone_8khz.flo", "white_noise.flo",
]


def main() -> int:
    src = os.path.join(REF, "Examples")
    if not os.path.isdir(src):
        print("reference not present at", src)
        return 1
    dst = os.path.join(HERE, "examples")
    os.makedirs(dst, exist_ok=True)
    manifest = {}
    for name in LOSSLESS:
        shutil.copyfile(os.path.join(src, name), os.path.join(dst, name))
        data = open(os.path.join(dst, name), "rb").read()
        manifest[name] = {"bytes": len(data), "sha256": hashlib.sha256(data).hexdigest()}
    wav = open(os.path.join(src, "audio.wav"), "rb").read()
    with gzip.GzipFile(os.path.join(dst, "audio.wav.gz"), "wb", mtime=0) as g:
        g.write(wav)
    manifest["audio.wav"] = {"bytes": len(wav), "sha256": hashlib.sha256(wav).hexdigest()}
    json.dump(manifest, open(os.path.join(dst, "MANIFEST.json"), "w"), indent=1, sort_keys=True)
    print("wrote", len(manifest), "fixtures to", dst)
    return 0


if __name__ == "__main__":
    sys.exit(main())
