"""Image sink / source (gms_b200/io_image.py): the host encoder against PIL (CPU), and -- under -m gpu -- the whole sink
against the call it replaces, torchvision.utils.save_image (scripts/render_time_animated.py:86-87), pixel for pixel."""
import io
import os

import numpy as np
import pytest
import torch

from gms_b200 import io_image


def _pil():
    return pytest.importorskip("PIL.Image")


@pytest.mark.parametrize("shape", [(7, 5, 3), (64, 33, 3), (16, 16, 1), (9, 4, 4)])
def test_png_encoder_is_read_back_identically_by_pil(shape):
    Image = _pil()
    H, W, C = shape
    rs = np.random.RandomState(H)
    img = rs.randint(0, 256, size=shape, dtype=np.uint8)
    lines = np.concatenate([np.zeros((H, 1), np.uint8), img.reshape(H, W * C)], axis=1)
    data = io_image.encode_png(lines.tobytes(), W, H, C, level=1)
    back = np.asarray(Image.open(io.BytesIO(data)))
    np.testing.assert_array_equal(back.reshape(H, W, C), img)
    np.testing.assert_array_equal(io_image.decode_png(data), img)


def test_png_decoder_reads_foreign_files_with_every_filter_type(tmp_path):
    Image = _pil()
    rs = np.random.RandomState(0)
    smooth = (np.add.outer(np.arange(40), np.arange(56))[:, :, None] * np.array([1, 2, 3]) % 256).astype(np.uint8)   # makes PIL pick Sub/Up/Paeth
    noisy = rs.randint(0, 256, size=(40, 56, 3), dtype=np.uint8)
    for k, img in enumerate((smooth, noisy)):
        p = str(tmp_path / f"f{k}.png")
        Image.fromarray(img).save(p, optimize=True)
        np.testing.assert_array_equal(io_image.decode_png(open(p, "rb").read()), img)
        np.testing.assert_array_equal(io_image.load_image_u8(p).numpy(), img)


@pytest.mark.gpu
def test_sink_png_equals_torchvision_save_image(tmp_path):
    Image = _pil()
    tvu = pytest.importorskip("torchvision.utils")
    g = torch.Generator().manual_seed(0)
    H, W = 270, 481
    frames = [(torch.rand(3, H, W, generator=g) * 1.2 - 0.1).cuda() for _ in range(6)]
    frames[1][0, 0, :4] = torch.tensor([0.0, 1.0, 0.5 / 255, 254.5 / 255]).cuda()      # rounding edges
    with io_image.ImageSink(H, W, fmt="png", slots=2, workers=2) as sink:
        for k, f in enumerate(frames):
            sink.write(f, str(tmp_path / "ours" / f"{k:05d}.png"))
    for k, f in enumerate(frames):
        ref = str(tmp_path / f"ref{k}.png")
        tvu.save_image(f, ref)
        a = np.asarray(Image.open(str(tmp_path / "ours" / f"{k:05d}.png")))
        b = np.asarray(Image.open(ref))
        np.testing.assert_array_equal(a, b)


@pytest.mark.gpu
def test_sink_ppm_raw_and_u8_source_round_trip(tmp_path):
    g = torch.Generator().manual_seed(1)
    H, W = 64, 80
    frames = [torch.rand(3, H, W, generator=g).cuda() for _ in range(9)]
    want = [f.mul(255).add_(0.5).clamp_(0, 255).permute(1, 2, 0).to("cpu", torch.uint8).numpy() for f in frames]
    raw = str(tmp_path / "frames.rgb")
    with io_image.ImageSink(H, W, fmt="raw", raw_path=raw, slots=3, workers=3) as sink:
        for f in frames:
            sink.write(f)
    got = np.fromfile(raw, np.uint8).reshape(len(frames), H, W, 3)
    for k in range(len(frames)):
        np.testing.assert_array_equal(got[k], want[k])            # submission order kept with 3 encoder threads
    io_image.save_image(frames[0], str(tmp_path / "a.ppm"))
    u8 = io_image.load_image_u8(str(tmp_path / "a.ppm"))
    np.testing.assert_array_equal(u8.numpy(), want[0])
    back = io_image.to_device_float(u8.cuda(non_blocking=True), hwc=True)
    ref = torch.from_numpy(want[0]).permute(2, 0, 1).float().div(255)        # ToTensor
    assert torch.equal(back.cpu(), ref)
    back2 = io_image.to_device_float(torch.from_numpy(want[0]).permute(2, 0, 1).contiguous().cuda(), hwc=False)
    assert torch.equal(back2.cpu(), ref)
