"""Per-patch prediction writer (mirrors src/flair/writer.py): PRED_<image name>, uint8, 0-based class
ids, LZW; georeferenced output copies the input's GeoTIFF tags (writer.py:38-43), plain output is a
striped LZW TIFF like PIL's (writer.py:50)."""
from __future__ import annotations

from pathlib import Path

import numpy as np

from .. import geotiff


class predictionwriter:
    def __init__(self, config, output_dir, write_interval="batch"):
        self.config = config
        self.output_dir = str(output_dir)
        Path(self.output_dir).mkdir(exist_ok=True, parents=True)

    def _write_one(self, pred: np.ndarray, filename: str) -> None:
        name = filename.split("/")[-1]
        output_file = str(self.output_dir + "/" + "PRED_" + name)
        if self.config["georeferencing_output"]:
            tags = geotiff.read_info(filename).geo_tags
            geotiff.write(output_file, pred, geo_tags=tags, compress="lzw", tiled=False, blocksize=64, bigtiff=False)
        else:
            geotiff.write(output_file, pred, compress="lzw", tiled=False, blocksize=64, bigtiff=False)

    def write_on_batch_end(self, prediction: dict) -> None:
        preds = prediction["preds"].cpu().numpy().astype("uint8")
        for pred, filename in zip(preds, prediction["id"]):
            self._write_one(pred, filename)

    def write_async(self, prediction: dict, pool) -> list:
        """write_on_batch_end with one task per file on `pool` (the LZW encoder releases the GIL); returns the futures."""
        preds = prediction["preds"].cpu().numpy().astype("uint8")
        return [pool.submit(self._write_one, pred, filename) for pred, filename in zip(preds, prediction["id"])]
