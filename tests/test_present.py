"""Present step and image files: the reference blits the RGBA32F result image into the B8G8R8A8_SRGB
swapchain image right after the dispatch (RenderTechniquePT.cpp:143-167, VulkanSwapchain.cpp:68-79).  Headless, the same
conversion feeds a PNG writer; the linear image goes to PFM; the accumulation state can be checkpointed and resumed."""
import ctypes as C
import json
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "cloudrendering_vulkan_b200", "bin", "crv_headless")


def _read_png(path):
    """minimal PNG reader (8-bit RGBA, filter 0 rows): checks the signature and every chunk CRC"""
    b = open(path, "rb").read()
    assert b[:8] == b"\x89PNG\r\n\x1a\n"
    off, idat, w = 8, b"", None
    while off < len(b):
        n, typ = struct.unpack(">I4s", b[off:off + 8])
        data = b[off + 8:off + 8 + n]
        crc, = struct.unpack(">I", b[off + 8 + n:off + 12 + n])
        assert zlib.crc32(typ + data) == crc, f"bad CRC in chunk {typ}"
        if typ == b"IHDR":
            w, h, depth, ctype, comp, filt, inter = struct.unpack(">IIBBBBB", data)
            assert (depth, ctype, comp, filt, inter) == (8, 6, 0, 0, 0)
        elif typ == b"IDAT":
            idat += data
        off += 12 + n
    raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, w * 4 + 1)
    assert (raw[:, 0] == 0).all()
    return raw[:, 1:].reshape(h, w, 4)


# ---- CPU: the oracle's restatement of the format conversion, and the PNG writer --------------------------------------
def test_oracle_blit_known_answers(orc):
    px = np.array([[[0.0, 1.0, 0.5, 0.5], [0.2, 0.0031308, -3.0, 2.0], [np.nan, 0.0031309, 1e-4, 0.25], [7.0, 0.003, 0.999, 0.0]]],
                  np.float32)
    out = orc.blit_srgb8(px)
    # bytes are B, G, R, A; colour through the sRGB transfer function, alpha linear; clamp; NaN -> 0
    assert out[0, 0].tolist() == [188, 255, 0, 128]
    assert out[0, 1].tolist() == [0, 10, 124, 255]
    assert out[0, 2].tolist() == [0, 10, 0, 64]
    assert out[0, 3].tolist() == [255, 10, 255, 0]
    # monotone and onto: a ramp covers every code
    ramp = np.zeros((1, 4096, 4), np.float32)
    ramp[0, :, 0] = np.linspace(0, 1, 4096, dtype=np.float32)
    r = orc.blit_srgb8(ramp)[0, :, 2].astype(int)
    assert (np.diff(r) >= 0).all() and set(r.tolist()) == set(range(256))


def test_png_writer_roundtrip(crv, tmp_path):
    rng = np.random.default_rng(5)
    for (w, h) in [(1, 1), (7, 3), (300, 70)]:             # the last one spans more than one stored deflate block
        px = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        path = str(tmp_path / f"t_{w}x{h}.png")
        crv.capi.check(crv.capi.lib.crv_write_png_rgba8(path.encode(), w, h, px.ctypes.data_as(C.c_void_p)))
        assert np.array_equal(_read_png(path), px)
    assert crv.capi.lib.crv_write_png_rgba8(b"/nonexistent_dir/x.png", 1, 1, px.ctypes.data_as(C.c_void_p)) == crv.capi.CRV_ERR_IO


# ---- GPU ----------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_blit_srgb8_matches_the_oracle(crv, orc, ctx):
    rng = np.random.default_rng(9)
    w, h = 257, 63
    px = rng.random((h, w, 4), dtype=np.float32)
    px[0, :8] = [[-1, 0, 1, 2]] * 8
    px[1, :4, 0] = [np.nan, 0.0031308, 0.0031309, 1e-30]
    px[2] *= 0.01                                           # the linear segment
    img = crv.Image(ctx, w, h)
    img.upload(px)
    gpu = img.blit_srgb8()
    cpu = orc.blit_srgb8(px)
    d = np.abs(gpu.astype(int) - cpu.astype(int))
    print(f"blit: {(d != 0).mean():.2e} of the bytes differ (powf of the two math libraries), max {d.max()} LSB")
    assert d.max() <= 1 and (d != 0).mean() <= 1e-3
    assert np.array_equal(gpu[..., 3], cpu[..., 3])        # alpha has no transcendental: exact
    img.close()


@pytest.mark.gpu
def test_png_pfm_and_checkpoint_files(crv, orc, ctx, tmp_path):
    from tests.test_gpu_parity import make_scene, _bits
    r, _ = make_scene(crv, orc, ctx, layout="brick", g=0.85, absorption=0.1)
    r.render(3)
    img = r.resultImage
    png, pfm, ck = (str(tmp_path / n) for n in ("a.png", "a.pfm", "a.ckpt"))
    img.save_png(png)
    bgra = img.blit_srgb8()
    assert np.array_equal(_read_png(png), bgra[..., [2, 1, 0, 3]])
    img.save_pfm(pfm)
    raw = open(pfm, "rb").read()
    hdr = f"PF\n{img.width} {img.height}\n-1.0\n".encode()
    assert raw.startswith(hdr)
    rgb = np.frombuffer(raw[len(hdr):], "<f4").reshape(img.height, img.width, 3)[::-1]
    lin = img.download()
    assert np.array_equal(_bits(rgb), _bits(lin[..., :3]))
    # checkpoint after 3 frames, 2 more -> A; fresh renderer resumed from the file, the same 2 frames -> identical bits
    img.save_checkpoint(ck, r.pushConstants.frameCount, 21238)
    r.render(2)
    a = img.download().copy()
    r.close()
    r2, _ = make_scene(crv, orc, ctx, layout="brick", g=0.85, absorption=0.1)
    fc, seed = r2.resultImage.load_checkpoint(ck)
    assert (fc, seed) == (3, 21238)
    r2.pushConstants.frameCount = fc
    r2._frames_rendered = fc
    r2.render(2)
    assert np.array_equal(_bits(a), _bits(r2.resultImage.download()))
    with pytest.raises(crv.CrvError):
        crv.Image(ctx, 8, 8).load_checkpoint(ck)           # size mismatch is an error, not a silent resize
    r2.close()


@pytest.mark.gpu
def test_cpp_driver_resume_equals_uninterrupted_run(tmp_path):
    def run(args):
        r = subprocess.run([BIN] + [str(a) for a in args], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        return json.loads(r.stdout)
    base = ["--fbm", 32, "--width", 64, "--height", 48, "--nsv", 40, "--g", 0.85, "--absorption", 0.1]
    ck, png = str(tmp_path / "acc.ckpt"), str(tmp_path / "out.png")
    whole = run(base + ["--spp", 7])
    run(base + ["--spp", 4, "--checkpoint", ck])
    resumed = run(base + ["--spp", 3, "--resume", ck, "--png", png])
    assert resumed["image_fnv1a"] == whole["image_fnv1a"]
    assert _read_png(png).shape == (48, 64, 4)
