"""CPU oracle for the zone_detect / patch-predict hot path of Draghoyns/FLAIR-1.

TEST INFRASTRUCTURE ONLY. Nothing under flair-1_b200/ imports this package; only tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline / `--impl reference` legs may, and only as the
checker or as the timed CPU baseline -- never as a fallback for the CUDA path.

It restates, function by function and citing file:line, what the reference computes on this path
(plain PyTorch fp32 for the network, numpy for integer/byte work, sklearn for the confusion matrix).

Parity pinning status
---------------------
The reference ships no tests, golden vectors or fixtures for this path and cannot be
imported here (segmentation_models_pytorch, rasterio, geopandas, shapely, pytorch_lightning,
skimage missing; no network). What can be pinned is pinned: tests/golden/make_golden.py executes the
reference's OWN source for every pure-Python function on the path (slice_extent with stubbed I/O,
get_stride, patch_weights, total_weights, convert, norm, the metric formulas, clean_confmat,
parsing_metadata, get_module, load_checkpoint) straight from /root/reference and stores inputs and
outputs under tests/golden/; tests/test_oracle_golden.py checks this oracle against those vectors.
The network itself (segmentation-models-pytorch==0.3.3 `Unet("resnet34")`, setup.py:36, not vendored)
is restated from its published architecture and pinned only by its state_dict key set (278 entries),
parameter count (24,438,399 for 3 bands / 15 classes = README.md:91 "about 24.4M") and output shape:
for the forward pass itself **parity is unpinned** against real smp / the shipped .pth (absent,
.MISSING_LARGE_BLOBS:3).
"""
