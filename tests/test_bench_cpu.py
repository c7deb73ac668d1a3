"""bench.py's pure host logic: the algorithmic FLOP counts behind `roofline.achieved`, the clock-sample
parser behind `clocks`, and the CPU reference arm's JSON contract."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_algorithmic_flops_match_the_survey_table():
    fwd, bwd = bench.trunk_flops_per_image(300, 512)          # 380.01 GF forward, 758.28 GF backward per 512x512 image
    assert abs(fwd / 1e9 - 380.01) < 0.05 and abs(bwd / 1e9 - 758.28) < 0.1
    fwd, bwd = bench.trunk_flops_per_image(1024, 512)
    assert abs(fwd / 1e9 - 381.72) < 0.05 and abs(bwd / 1e9 - 761.71) < 0.1
    fwd, bwd = bench.trunk_flops_per_image(20, 512)
    assert abs(fwd / 1e9 - 379.35) < 0.05 and abs(bwd / 1e9 - 756.95) < 0.1


def test_conv_flops_reads_the_abi_argument_positions():
    """conv_flops slices (B, H, W, Cin, Cout, R, S, pad) out of the C-ABI argument tuples: positions per include/szn.h."""
    # szn_conv_fwd(dtype, x, wt, bias, y, B, H, W, Cin, Cout, R, S, pad, relu, scale, scale_ld, out_fp32, ldo, stream)
    a = (0, 1, 2, 3, 4, 8, 178, 178, 256, 256, 3, 3, 1, 1, None, 0, 0, 256, 0)
    want = 2.0 * 8 * 178 * 178 * 256 * 9 * 256
    assert bench.conv_flops("szn_conv_fwd", a) == want
    # szn_conv_dgrad(dtype, dy, wt_dgrad, dx, B, H, W, Cin, Cout, R, S, pad, relu_ref, scale, scale_ld, ld_dy, col_sum, stream)
    a = (0, 1, 2, 3, 8, 178, 178, 256, 256, 3, 3, 1, None, None, 0, 256, None, 0)
    assert bench.conv_flops("szn_conv_dgrad", a) == want
    # szn_conv_wgrad(dtype, x, dy, dw, B, H, W, Cin, Cout, R, S, pad, ld_dy, stream)
    a = (0, 1, 2, 3, 8, 178, 178, 256, 256, 3, 3, 1, 256, 0)
    assert bench.conv_flops("szn_conv_wgrad", a) == want
    # fc6: 7x7 valid on 23x23 -> 17x17 outputs
    a = (0, 1, 2, 3, 4, 8, 23, 23, 512, 4096, 7, 7, 0, 1, None, 0, 0, 4096, 0)
    assert bench.conv_flops("szn_conv_fwd", a) == 2.0 * 8 * 17 * 17 * 4096 * 49 * 512
    assert bench.conv_flops("szn_pool_fwd", a) == 0


def test_hbm_bytes_reads_the_abi_argument_positions():
    """hbm_bytes: algorithmic bytes of the HBM-bound kernels from their C-ABI argument tuples (include/szn.h)."""
    B, D, H = 8, 300, 512
    # szn_embed_loss_fwd(kind, score, target, target_embed, table, table_rows, n, c, h, w, stats, accum, loss, stream)
    a = (0, 1, 2, None, 4, 59, B, D, H, H, 5, 6, 7, 0)
    assert bench.hbm_bytes("szn_embed_loss_fwd", a, 4) == B * H * H * (D * 4 + 8)
    # szn_embed_argmax(score, table, n, D, h, w, C, scratch, labels, stream)
    assert bench.hbm_bytes("szn_embed_argmax", (1, 2, B, D, H, H, 59, 3, 4, 0), 4) == B * H * H * (D * 4 + 8)
    # szn_upsample32_crop_fwd(s, out, B, D, H, W, hs, ws, ld, coff, stream) / _bwd(dtype, g, ds, B, D, H, W, ...)
    assert bench.hbm_bytes("szn_upsample32_crop_fwd", (1, 2, B, D, H, H, 17, 17, 320, 0, 0), 4) == B * D * H * H * 4
    assert bench.hbm_bytes("szn_upsample32_crop_bwd", (0, 1, 2, B, D, H, H, 17, 17, 320, 0, 0), 4) == B * D * H * H * 4
    # szn_conv1_1_fwd(dtype, x, w, bias, y, B, H, W, pad, stream): image in, 710 x 710 x 64 out
    assert bench.hbm_bytes("szn_conv1_1_fwd", (0, 1, 2, 3, 4, B, H, H, 100, 0), 4) == B * 3 * H * H * 4 + B * 710 * 710 * 64 * 4
    # szn_pool_bwd(dtype, y, dp, dy, B, H, W, C, relu_gate, col_sum, stream)
    assert bench.hbm_bytes("szn_pool_bwd", (1, 1, 2, 3, B, 710, 710, 64, 1, None, 0), 2) == B * 710 * 710 * 64 * 2 * 2.25
    assert bench.hbm_bytes("szn_conv_fwd", a, 4) == 0


def test_every_baseline_config_is_selectable():
    assert sorted(bench.CONFIGS) == [0, 1, 2, 3, 4]
    assert bench.CONFIGS[1]["B"] == 8 and bench.CONFIGS[1]["precision"] == "tf32" and bench.CONFIGS[2]["B"] == 32
    assert bench.CONFIGS[3].get("zeroshot") and bench.CONFIGS[4]["D"] == 1024 and bench.CONFIGS[4]["C"] == 256
    assert len(bench.VAL_UNSEEN) == 10 and set(bench.VAL_UNSEEN).isdisjoint(bench.TRAIN_UNSEEN)


def test_clock_sample_parser(tmp_path):
    p = tmp_path / "clk.csv"
    p.write_text("0, 1965, 1965, 310.2, 0x0, Not Active, Not Active, Not Active, Not Active\n"
                 "0, 1890, 1965, 998.1, 0x4, Not Active, Not Active, Not Active, Active\n"
                 "0, 1800, 1965, 999.0, 0x4, Not Active, Not Active, Not Active, Active\n"
                 "1, 300, 1965, 80.0, 0x1, Not Active, Not Active, Not Active, Not Active\n"
                 "garbage line\n")
    c = bench.clocks_summary(str(p), 0)
    assert c == {"sm_mhz": 1890.0, "sm_max_mhz": 1965.0, "reasons": ["sw_power_cap"], "samples": 3}
    assert bench.clocks_summary(str(tmp_path / "missing.csv"), 0)["samples"] == 0


@pytest.mark.slow
def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` on the host cores (the CPU oracle port; one 512x512 image per step)."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["unit"] == "Mpixel/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["config"]["workload"].startswith("configs[1]")
