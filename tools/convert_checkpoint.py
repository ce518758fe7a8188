#!/usr/bin/env python
"""Convert a MobileSAM checkpoint (`mobile_sam.pt`, the file `script/export_models.py:21-43` of the reference
exports the three .onnx graphs from) into the engine's weight container.

    python tools/convert_checkpoint.py /path/to/mobile_sam.pt <model_directory>

writes <model_directory>/segmentation/mobile_sam_b200.bin.  The container keeps the checkpoint's tensor names
(SURVEY Appendix A.7), so the conversion is a plain dump of the floating-point entries; BatchNorm folding and all
re-layouts happen when the engine loads the file.  Not exercisable offline (no checkpoint on this box): the name
and shape contract is enforced by the loader (csrc/model.cu) and, for synthetic weights, by the oracle's strict load.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    if len(sys.argv) != 3:
        print(__doc__)
        sys.exit(2)
    import torch
    from dlimgedit_b200 import weights_io
    sd = torch.load(sys.argv[1], map_location="cpu")
    if isinstance(sd, dict) and "model" in sd and isinstance(sd["model"], dict):
        sd = sd["model"]
    tensors = weights_io.from_state_dict(sd)
    needed = ("image_encoder.patch_embed.seq.0.c.weight", "prompt_encoder.no_mask_embed.weight",
              "mask_decoder.iou_token.weight")
    missing = [k for k in needed if k not in tensors]
    if missing:
        sys.exit(f"this does not look like a MobileSAM (vit_t) checkpoint, missing {missing}")
    out_dir = os.path.join(sys.argv[2], "segmentation")
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, weights_io.WEIGHT_FILE_NAME)
    weights_io.save(path, tensors)
    print(f"wrote {path}: {len(tensors)} tensors, {sum(v.size for v in tensors.values())} parameters")


if __name__ == "__main__":
    main()
