"""Host-side planning of the tensor-core conv kernel (csrc/szn_umma.cu: pick_tile / pick_block_n / fill_sms), replayed in
Python for the layers of one 512x512 step: M-tile padding (pixel slots computed vs pixels needed) and wave quantisation on
148 SMs.  `--full` lets the tile search try every width instead of powers of two / exact widths (SZN_TILE_SEARCH=full).
No GPU needed.    python tools/plan_analysis.py [--batch 8] [--full]"""
import argparse
import math

LAYERS = [("conv1_2", 710, 64, 64, 3), ("conv2_1", 355, 64, 128, 3), ("conv2_2", 355, 128, 128, 3),
          ("conv3_1", 178, 128, 256, 3), ("conv3_2", 178, 256, 256, 3), ("conv3_3", 178, 256, 256, 3),
          ("conv4_1", 89, 256, 512, 3), ("conv4_2", 89, 512, 512, 3), ("conv4_3", 89, 512, 512, 3),
          ("conv5_1", 45, 512, 512, 3), ("conv5_2", 45, 512, 512, 3), ("conv5_3", 45, 512, 512, 3)]


def pick_tile(W, H, max_rows=128, full=False):
    best, bw, bh = -1, max_rows, 1
    for tw in range(1, min(max_rows, 256) + 1):
        th = min(max_rows // tw, 256)
        if th < 1:
            break
        pow2 = tw & (tw - 1) == 0
        if not full and not pow2 and tw != W and tw != (W + 1) // 2:
            continue
        cnt = math.ceil(W / tw) * math.ceil(H / th)
        if best < 0 or cnt < best or (cnt == best and tw > bw):
            best, bw, bh = cnt, tw, th
    return bw, min(bh, H)


def plan(hw, cout, B, full):
    tw, th = pick_tile(hw, hw, 128, full)
    m_tiles = math.ceil(hw / tw) * math.ceil(hw / th) * B
    block_n = 256 if cout >= 256 else cout
    while block_n > 64 and m_tiles * math.ceil(cout / block_n) < 148:
        block_n //= 2
    tiles = m_tiles * math.ceil(cout / block_n)
    return tw, th, block_n, tiles, m_tiles * 128 / (B * hw * hw), tiles / (math.ceil(tiles / 148) * 148)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--full", action="store_true")
    a = ap.parse_args()
    print("%-8s %5s | %9s %7s %7s | %8s %8s %8s" % ("layer", "HxW", "tile", "block_n", "tiles", "padding", "waves", "useful"))
    tot_w = tot = 0.0
    for name, hw, cin, cout, k in LAYERS:
        tw, th, bn, tiles, pad, wave = plan(hw, cout, a.batch, a.full)
        flops = hw * hw * cin * cout * k * k
        useful = wave / pad
        tot_w += flops / useful
        tot += flops
        print("%-8s %5d | %4dx%-4d %7d %7d | %7.1f%% %7.1f%% %7.1f%%" % (name, hw, tw, th, bn, tiles, (pad - 1) * 100, wave * 100,
                                                                       useful * 100))
    print("FLOP-weighted useful fraction of the issued MMA work (forward / dgrad of these layers): %.1f%%" % (tot / tot_w * 100))


if __name__ == "__main__":
    main()
