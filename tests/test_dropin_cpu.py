"""The UNMODIFIED reference training script on top of this package (build container only: needs
/root/reference).  On a machine without a GPU the run must get through the reference's own argument
parsing, dataset, model / criterion / optimizer / scheduler construction and into the first training
step, and stop there with OUR loud "no CPU path" error — proving the flat-import drop-in wiring
 without any CPU fallback existing."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF_TRAIN = "/root/reference/src/train.py"


def _drop_plot_stubs():
    """The launcher's stand-ins for matplotlib / seaborn sub-modules must not outlive a test."""
    for k in [k for k in sys.modules if k.split(".")[0] in ("matplotlib", "seaborn")]:
        if getattr(sys.modules[k], "__file__", None) is None:
            del sys.modules[k]


def _make_mvtec_tree(root, category="synth", n_train=4, n_test=2, size=40):
    from PIL import Image
    rng = np.random.default_rng(0)
    for split, n in (("train", n_train), ("test", n_test)):
        d = os.path.join(root, category, split, "good")
        os.makedirs(d, exist_ok=True)
        for i in range(n):
            Image.fromarray(rng.integers(0, 255, (size, size, 3), dtype=np.uint8)).save(os.path.join(d, f"{i:03d}.png"))
    os.makedirs(os.path.join(root, category, "ground_truth"), exist_ok=True)


@pytest.mark.skipif(not os.path.exists(REF_TRAIN), reason="reference sources only exist in the build container")
def test_reference_train_script_runs_on_top_of_this_package(tmp_path, monkeypatch):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU wiring test")
    pytest.importorskip("torchvision")
    from tiaozhanbei_unet_b200 import _lib
    from tiaozhanbei_unet_b200.dropin import launch
    data = str(tmp_path / "mvtec")
    _make_mvtec_tree(data)
    saved_modules = {k: sys.modules.get(k) for k in ("model", "train_utils", "utils", "dataset", "matplotlib", "seaborn")}
    saved_argv, saved_path = list(sys.argv), list(sys.path)
    monkeypatch.chdir(os.path.dirname(REF_TRAIN))
    try:
        with pytest.raises((_lib.UnetkError, RuntimeError)) as e:
            launch.run(REF_TRAIN, ["--data_root", data, "--category", "synth", "--image_size", "32", "--epochs", "1",
                                   "--batch_size", "2", "--num_workers", "0", "--use_ssim", "--model", "anomaly_unet",
                                   "--save_dir", str(tmp_path / "out"), "--device", "cpu"])
        assert "CUDA" in str(e.value) or "CPU" in str(e.value)
        # the script really used OUR modules under the reference's flat names
        import model as flat_model
        import train_utils as flat_tu
        assert flat_model.AnomalyUNet.__module__ == "tiaozhanbei_unet_b200.model"
        assert flat_tu.CombinedLoss.__module__ == "tiaozhanbei_unet_b200.train_utils"
        assert os.path.isdir(str(tmp_path / "out"))      # experiment directory + args.json were written first
    finally:
        sys.argv[:] = saved_argv
        sys.path[:] = saved_path
        for k, v in saved_modules.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        _drop_plot_stubs()


REF_SEG_TRAIN = "/root/reference/train.py"


def _make_gear_tree(root, n=4, size=48):
    """images/<split>/*.png + labels/<split>/*.txt, one normalised polygon per line (src/gear_dataset.py:66-150)."""
    from PIL import Image
    rng = np.random.default_rng(1)
    for split in ("train", "val", "test"):
        os.makedirs(os.path.join(root, "images", split), exist_ok=True)
        os.makedirs(os.path.join(root, "labels", split), exist_ok=True)
        for i in range(n):
            Image.fromarray(rng.integers(0, 255, (size, size, 3), dtype=np.uint8)).save(
                os.path.join(root, "images", split, f"{i:03d}.png"))
            with open(os.path.join(root, "labels", split, f"{i:03d}.txt"), "w") as f:
                f.write("0 0.1 0.1 0.5 0.1 0.5 0.5 0.1 0.5\n")      # pitting
                f.write("1 0.4 0.4 0.9 0.4 0.9 0.9\n")              # spalling (overlaps: priority resolution)
                f.write("2 0.6 0.1 0.9 0.1 0.9 0.3 0.6 0.3\n")      # scrape


@pytest.mark.skipif(not os.path.exists(REF_SEG_TRAIN), reason="reference sources only exist in the build container")
def test_reference_segmentation_train_script_runs_on_top_of_this_package(tmp_path, monkeypatch):
    """Root train.py (Gear pipeline, train.py:20-23 imports src.model / src.metrics): dataset, 4-class
    SegmentationUNet, CombinedSegmentationLoss and Adam are built from OUR modules, then the first step stops
    at the loud no-CPU-path error."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU wiring test")
    pytest.importorskip("torchvision")
    pytest.importorskip("sklearn")
    from tiaozhanbei_unet_b200 import _lib
    from tiaozhanbei_unet_b200.dropin import launch
    data = str(tmp_path / "Gear")
    _make_gear_tree(data)
    names = ("model", "train_utils", "matplotlib", "seaborn", "src", "src.model", "src.metrics", "src.utils",
             "src.gear_dataset")
    saved_modules = {k: sys.modules.get(k) for k in names}
    saved_argv, saved_path = list(sys.argv), list(sys.path)
    monkeypatch.chdir(os.path.dirname(REF_SEG_TRAIN))
    try:
        with pytest.raises((_lib.UnetkError, RuntimeError)) as e:
            launch.run(REF_SEG_TRAIN, ["--data_root", data, "--image_size", "32", "--epochs", "1", "--batch_size", "2",
                                       "--num_workers", "0", "--model", "seg_unet", "--focal_weight", "0.5",
                                       "--class_weights", "0.5,1,1,2", "--save_dir", str(tmp_path / "out"),
                                       "--device", "cpu"])
        assert "CUDA" in str(e.value) or "CPU" in str(e.value)
        import src.metrics as sm
        import src.model as smodel
        assert smodel.SegmentationUNet.__module__ == "tiaozhanbei_unet_b200.model"
        assert sm.CombinedSegmentationLoss.__module__ == "tiaozhanbei_unet_b200.metrics"
        assert sm.dice_loss.__module__ == "tiaozhanbei_unet_b200.metrics"
        assert sm.SegmentationMetrics.__module__ == "src.metrics"      # reporting half stays the reference's ...
        from tiaozhanbei_unet_b200.metrics import DeviceConfusion
        calc = sm.SegmentationMetrics(4)
        assert isinstance(calc, DeviceConfusion)                        # ... over the device-side confusion counts
        assert type(calc).compute_all_metrics.__module__ == "src.metrics" and type(calc).update is DeviceConfusion.update
        assert calc.confusion_matrix.shape == (4, 4) and calc.compute_all_metrics()["mean_iou"] == 0.0
        assert os.path.isdir(str(tmp_path / "out"))
    finally:
        sys.argv[:] = saved_argv
        sys.path[:] = saved_path
        for k, v in saved_modules.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        _drop_plot_stubs()


REF_KOLEKTOR_TRAIN = "/root/reference/train_kolektorsdd.py"


def _make_kolektor_tree(root, folders=4, per_folder=3, h=60, w=30):
    """kosNN/PartK.jpg + PartK_label.bmp (src/kolektorsdd_dataset.py:44-58)."""
    from PIL import Image
    rng = np.random.default_rng(2)
    for f in range(folders):
        d = os.path.join(root, f"kos{f + 1:02d}")
        os.makedirs(d, exist_ok=True)
        for k in range(per_folder):
            Image.fromarray(rng.integers(0, 255, (h, w), dtype=np.uint8)).save(os.path.join(d, f"Part{k}.jpg"))
            label = np.zeros((h, w), dtype=np.uint8)
            label[10:20, 5:15] = 1 + (k % 2)
            Image.fromarray(label).save(os.path.join(d, f"Part{k}_label.bmp"))


@pytest.mark.skipif(not os.path.exists(REF_KOLEKTOR_TRAIN), reason="reference sources only exist in the build container")
def test_reference_kolektor_train_script_runs_on_top_of_this_package(tmp_path, monkeypatch):
    """train_kolektorsdd.py (3 classes, 2:1 strips, class weights 1/50/50 by default, bilinear decoders here): same
    wiring as the Gear script, stops at the loud no-CPU-path error of the first step."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU wiring test")
    pytest.importorskip("torchvision")
    pytest.importorskip("sklearn")
    from tiaozhanbei_unet_b200 import _lib
    from tiaozhanbei_unet_b200.dropin import launch
    data = str(tmp_path / "KolektorSDD")
    _make_kolektor_tree(data)
    names = ("model", "train_utils", "matplotlib", "seaborn", "src", "src.model", "src.metrics", "src.utils",
             "src.kolektorsdd_dataset")
    saved_modules = {k: sys.modules.get(k) for k in names}
    saved_argv, saved_path = list(sys.argv), list(sys.path)
    monkeypatch.chdir(os.path.dirname(REF_KOLEKTOR_TRAIN))
    try:
        with pytest.raises((_lib.UnetkError, RuntimeError)) as e:
            launch.run(REF_KOLEKTOR_TRAIN, ["--data_root", data, "--image_height", "64", "--image_width", "32",
                                            "--epochs", "1", "--batch_size", "2", "--num_workers", "0", "--bilinear",
                                            "--save_dir", str(tmp_path / "out"), "--device", "cpu"])
        assert "CUDA" in str(e.value) or "CPU" in str(e.value)
        import src.metrics as sm
        assert sm.CombinedSegmentationLoss.__module__ == "tiaozhanbei_unet_b200.metrics"
    finally:
        sys.argv[:] = saved_argv
        sys.path[:] = saved_path
        for k, v in saved_modules.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        _drop_plot_stubs()


REF_TEST = "/root/reference/src/test.py"


@pytest.mark.skipif(not os.path.exists(REF_TEST), reason="reference sources only exist in the build container")
def test_reference_test_script_loads_our_checkpoint(tmp_path, monkeypatch, capsys):
    """src/test.py: dataset with anomalous test images + masks, AnomalyUNet from OUR module, a checkpoint in the
    reference's dict layout (src/utils.py:37-58) loaded by the reference's own load_checkpoint with strict keys,
    then test_model stops at the loud no-CPU-path error of the first forward."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU wiring test")
    pytest.importorskip("torchvision")
    from PIL import Image
    from tiaozhanbei_unet_b200 import _lib
    from tiaozhanbei_unet_b200.dropin import launch
    from tiaozhanbei_unet_b200.model import AnomalyUNet
    data = str(tmp_path / "mvtec")
    _make_mvtec_tree(data)
    rng = np.random.default_rng(3)
    for sub, mode in (("test/scratch", "rgb"), ("ground_truth/scratch", "mask")):
        d = os.path.join(data, "synth", sub)
        os.makedirs(d, exist_ok=True)
        for i in range(2):
            if mode == "rgb":
                Image.fromarray(rng.integers(0, 255, (40, 40, 3), dtype=np.uint8)).save(os.path.join(d, f"{i:03d}.png"))
            else:
                m = np.zeros((40, 40), dtype=np.uint8)
                m[5:15, 5:15] = 255
                Image.fromarray(m).save(os.path.join(d, f"{i:03d}_mask.png"))
    ckpt = str(tmp_path / "ckpt.pth")
    torch.manual_seed(0)
    torch.save({"epoch": 3, "model_state_dict": AnomalyUNet(3).state_dict(), "optimizer_state_dict": {}, "loss": 0.25}, ckpt)
    saved_modules = {k: sys.modules.get(k) for k in ("model", "train_utils", "utils", "dataset", "matplotlib", "seaborn")}
    saved_argv, saved_path = list(sys.argv), list(sys.path)
    monkeypatch.chdir(os.path.dirname(REF_TEST))
    try:
        with pytest.raises((_lib.UnetkError, RuntimeError)) as e:
            launch.run(REF_TEST, ["--data_root", data, "--category", "synth", "--image_size", "32", "--batch_size", "2",
                                  "--num_workers", "0", "--model", "anomaly_unet", "--checkpoint", ckpt,
                                  "--threshold", "0.5", "--output_dir", str(tmp_path / "res"), "--device", "cpu"])
        assert "CUDA" in str(e.value) or "CPU" in str(e.value)
        out = capsys.readouterr().out
        assert "Test samples: 4" in out and "epoch 3, loss 0.2500" in out and "Testing model..." in out
    finally:
        sys.argv[:] = saved_argv
        sys.path[:] = saved_path
        for k, v in saved_modules.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        _drop_plot_stubs()


REF_SEG_TEST = "/root/reference/test.py"


@pytest.mark.skipif(not os.path.exists(REF_SEG_TEST), reason="reference sources only exist in the build container")
def test_reference_segmentation_test_script_loads_our_checkpoint(tmp_path, monkeypatch, capsys):
    """Root test.py (Gear evaluation): 4-class SegmentationUNet from OUR module + a checkpoint in the reference's dict
    layout through the reference's load_checkpoint, then evaluate_model stops at the loud no-CPU-path error."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU wiring test")
    pytest.importorskip("torchvision")
    pytest.importorskip("sklearn")
    from tiaozhanbei_unet_b200 import _lib
    from tiaozhanbei_unet_b200.dropin import launch
    from tiaozhanbei_unet_b200.model import SegmentationUNet
    data = str(tmp_path / "Gear")
    _make_gear_tree(data)
    ckpt = str(tmp_path / "seg.pth")
    torch.manual_seed(0)
    torch.save({"epoch": 7, "model_state_dict": SegmentationUNet(3, 4).state_dict(), "optimizer_state_dict": {}, "loss": 0.125}, ckpt)
    names = ("model", "train_utils", "matplotlib", "seaborn", "src", "src.model", "src.metrics", "src.utils",
             "src.gear_dataset")
    saved_modules = {k: sys.modules.get(k) for k in names}
    saved_argv, saved_path = list(sys.argv), list(sys.path)
    monkeypatch.chdir(os.path.dirname(REF_SEG_TEST))
    try:
        with pytest.raises((_lib.UnetkError, RuntimeError)) as e:
            launch.run(REF_SEG_TEST, ["--data_root", data, "--image_size", "32", "--batch_size", "2", "--num_workers", "0",
                                      "--model", "seg_unet", "--checkpoint", ckpt, "--save_dir", str(tmp_path / "res"),
                                      "--device", "cpu"])
        assert "CUDA" in str(e.value) or "CPU" in str(e.value)
        out = capsys.readouterr().out
        assert "epoch 7" in out and "0.1250" in out and "Model parameters: 31,0" in out
    finally:
        sys.argv[:] = saved_argv
        sys.path[:] = saved_path
        for k, v in saved_modules.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        _drop_plot_stubs()
