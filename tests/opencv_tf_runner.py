"""Runs serialized TensorFlow GraphDefs through OpenCV's TensorFlow importer (cv2.dnn.readNetFromTensorflow) in a process of its own:
OpenCV links its own protobuf runtime, which must not share a process with the Python protobuf package the graphs are built with.

    python tests/opencv_tf_runner.py <dir>      # <dir>/manifest.json: [{"name": ...}], <dir>/<name>.pb, <name>_x.npy -> <name>_y.npy

Blobs are NCHW inside OpenCV; the importer transposes NHWC graphs itself, so inputs / outputs are transposed here."""
import json
import os
import sys

import numpy as np
import cv2


def main(d):
    done = []
    for case in json.load(open(os.path.join(d, "manifest.json"))):
        name = case["name"]
        x = np.load(os.path.join(d, name + "_x.npy"))
        net = cv2.dnn.readNetFromTensorflow(os.path.join(d, name + ".pb"))
        net.setInput(np.ascontiguousarray(x.transpose(0, 3, 1, 2)) if x.ndim == 4 else x)
        y = net.forward()
        np.save(os.path.join(d, name + "_y.npy"), y.transpose(0, 2, 3, 1) if y.ndim == 4 else y)
        done.append(name)
    json.dump({"opencv": cv2.__version__, "done": done}, open(os.path.join(d, "done.json"), "w"))


if __name__ == "__main__":
    main(sys.argv[1])
