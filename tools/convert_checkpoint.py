#!/usr/bin/env python
"""Convert MobileSAM weights into the engine's weight container.

    python tools/convert_checkpoint.py /path/to/mobile_sam.pt <model_directory>
    python tools/convert_checkpoint.py --onnx-dir <dir with the reference's three .onnx files> <model_directory>

writes <model_directory>/segmentation/mobile_sam_b200.bin.

* `mobile_sam.pt` is the checkpoint the reference's graphs are exported from (script/export_models.py:21-43).  The container
  keeps the checkpoint's tensor names, so the conversion is a plain dump of the floating-point entries;
  BatchNorm folding and all re-layouts happen when the engine loads the file.
* `--onnx-dir` reads `mobile_sam_image_encoder.onnx` + `sam_mask_decoder_{multi,single}.onnx` (the files
  models/segmentation/CMakeLists.txt:2-16 downloads) with dlimgedit_b200/onnx_import.py: no onnx / onnxruntime needed.
  MD5 sums are compared with the ones pinned by the reference and a mismatch is reported (not fatal: re-exports differ).

Neither a checkpoint nor the .onnx files exist offline; tests/test_onnx_import.py covers both routes on stand-ins.
"""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    args = sys.argv[1:]
    from dlimgedit_b200 import weights_io
    if len(args) == 3 and args[0] == "--onnx-dir":
        from dlimgedit_b200 import onnx_import
        src, model_dir = args[1], args[2]
        for name, md5 in onnx_import.MD5.items():
            path = os.path.join(src, name)
            if os.path.exists(path):
                got = hashlib.md5(open(path, "rb").read()).hexdigest().upper()
                if got != md5:
                    print(f"note: {name} has MD5 {got}, the reference pins {md5}", file=sys.stderr)
        try:
            tensors = onnx_import.state_from_onnx_dir(src)
        except (ValueError, FileNotFoundError) as e:
            sys.exit(f"could not read the MobileSAM graphs: {e}")
    elif len(args) == 2:
        import torch
        sd = torch.load(args[0], map_location="cpu")
        if isinstance(sd, dict) and "model" in sd and isinstance(sd["model"], dict):
            sd = sd["model"]
        tensors = weights_io.from_state_dict(sd)
        model_dir = args[1]
    else:
        print(__doc__)
        sys.exit(2)
    needed = ("image_encoder.patch_embed.seq.0.c.weight", "prompt_encoder.no_mask_embed.weight",
              "mask_decoder.iou_token.weight")
    missing = [k for k in needed if k not in tensors]
    if missing:
        sys.exit(f"this does not look like a MobileSAM (vit_t) checkpoint, missing {missing}")
    out_dir = os.path.join(model_dir, "segmentation")
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, weights_io.WEIGHT_FILE_NAME)
    weights_io.save(path, tensors)
    print(f"wrote {path}: {len(tensors)} tensors, {sum(v.size for v in tensors.values())} parameters")


if __name__ == "__main__":
    main()
