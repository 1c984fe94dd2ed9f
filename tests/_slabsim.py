"""CPU replay of the SLAB feed's address arithmetic (test infrastructure only; never imported by the product).

The slab kernel (ted-spad_b200/csrc/conv_slab.cu) is driven entirely by a host-built plan: a 5-D TMA box per
(tile, K stage), UMMA shared-memory descriptor fields, and a table of (A offset, B offset) pairs.  This module
replays exactly that plan on the CPU with the hardware semantics pinned on the B200 by tests/probes/umma_probe.cu
(profiles/r1_umma_probe.txt):

  * TMA tiled load: box elements land row-major (innermost first) in shared memory, out-of-bounds elements are
    zero, and with SWIZZLE_128B the 16-byte chunk index (byte bits 4-6) is XORed with byte bits 7-9;
  * UMMA K-major operand, SWIZZLE_128B: element (row, k) at start + (row>>3)*SBO + (row&7)*128 + 2k, the same
    XOR applied to the final address (descriptor base_offset = 0);
  * UMMA K-major operand, no swizzle: element (row, k) at start + (row>>3)*SBO + (k>>3)*LBO + (row&7)*16 + 2(k&7),
    where LBO may be 16 bytes (overlapping K-adjacent core matrices).

so that tiling, descriptor tables and the weight image are verified against F.conv3d without a GPU.  It also holds
the reference (numpy) packer of the weight image, against which the CUDA pack kernel is compared bit for bit.
"""
import numpy as np
import torch

from tedspad_b200 import _lib as L


def bf16_bits(t):
    """torch bf16 tensor -> numpy uint16 bit patterns (flat)."""
    return t.contiguous().view(torch.int16).cpu().numpy().view(np.uint16).reshape(-1)


def bits_to_f32(u16):
    return (u16.astype(np.uint32) << 16).view(np.float32)


# ---------------------------------------------------------------------------------- weight image
def pack_image(kind, w_std, n_tile, k_pad, cin_pad, k, pw_front):
    """numpy restatement of slab_pack_kernel: w_std uint16 [Cout_pad*K_pad] -> image uint16."""
    kd, kh, kw = k
    w = w_std.reshape(n_tile, k_pad)
    if kind in (L.SLAB_3X3_PAIR, L.SLAB_STEM3D_PAIR):   # two single-CTA images of n_tile / 2 rows each, the leader CTA's first
        half = n_tile // 2
        single = L.SLAB_3X3 if kind == L.SLAB_3X3_PAIR else L.SLAB_STEM3D
        return np.concatenate([pack_image(single, w[h * half:(h + 1) * half].reshape(-1), half, k_pad, cin_pad, k, pw_front)
                               for h in (0, 1)])
    if kind == L.SLAB_3X3_KX_PAIR:
        # per CTA half: [ky*CB + cb] blocks of 3*CP/2 rows x 128 B; GEMM column = kx*CP + output channel
        cp, rows, cb_n = n_tile, 3 * n_tile // 2, cin_pad // 64
        halves = []
        for h in (0, 1):
            total = 3 * cb_n * rows * 128 // 2
            idx = np.arange(total, dtype=np.int64)
            byte = idx * 2
            blk_bytes = rows * 128
            blk, o = byte // blk_bytes, byte % blk_bytes
            grp, row, chunk_sw, within = o >> 10, (o >> 7) & 7, (o >> 4) & 7, (o & 15) >> 1
            col = h * rows + grp * 8 + row
            kx, n = col // cp, col % cp
            kk = ((chunk_sw ^ row) << 3) + within
            ky, cb = blk // cb_n, blk % cb_n
            halves.append(w[n, (ky * 3 + kx) * cin_pad + cb * 64 + kk])
        return np.concatenate(halves)
    if kind == L.SLAB_3X3:
        total = 9 * (cin_pad // 64) * n_tile * 128 // 2
        idx = np.arange(total, dtype=np.int64)
        byte = idx * 2
        blk_bytes = n_tile * 128
        blk, o = byte // blk_bytes, byte % blk_bytes
        grp, row, chunk_sw, within = o >> 10, (o >> 7) & 7, (o >> 4) & 7, (o & 15) >> 1
        n = grp * 8 + row
        kk = ((chunk_sw ^ row) << 3) + within
        cb_n = cin_pad // 64
        tap, cb = blk // cb_n, blk % cb_n
        src = tap * cin_pad + cb * 64 + kk
        return w[n, src]
    n_mma = 6 if kind == L.SLAB_STEM2D else kd * kh * 2
    total = n_mma * 2 * n_tile * 8
    idx = np.arange(total, dtype=np.int64)
    e = idx & 7
    t = idx >> 3
    n = t % n_tile
    t //= n_tile
    j = t & 1
    mma = t >> 1
    if kind == L.SLAB_STEM2D:
        ky, q = mma >> 1, mma & 1
        kx = 2 * q + j
        ok = (kx < 3) & (e < cin_pad)
        src = (ky * 3 + kx) * cin_pad + e
    else:
        shift = pw_front & 1
        q = mma & 1
        kyt = mma >> 1
        ky, kt = kyt % kh, kyt // kh
        px = 2 * (2 * q + j) + (e >> 2)
        ch = e & 3
        kx = px - shift
        ok = (kx >= 0) & (kx < kw) & (ch < cin_pad)
        src = ((kt * kh + ky) * kw + kx) * cin_pad + ch
    src = np.where(ok, src, 0)
    return np.where(ok, w[n, src], np.uint16(0))


# ------------------------------------------------------------------------------------- hardware
def tma_box(xbits, plan, coords, estr=None):
    """Emulate one 5-D tiled TMA load: returns the shared-memory image (uint16, slab_stride bytes, zero padded).
    estr: cuTensorMapEncodeTiled elementStrides (innermost first) - element i of the box along dimension d is tensor
    element coords[d] + i * estr[d] (plan.box counts the elements KEPT, as conv_slab.cu's forward passes them)."""
    box = list(plan.box)
    dims = list(plan.tdim)
    strides = [2] + list(plan.tstride)  # bytes
    ii = np.meshgrid(*[np.arange(b) for b in reversed(box)], indexing="ij")  # order: i4,i3,i2,i1,i0
    ii = [a.reshape(-1) for a in reversed(ii)]                                # back to i0..i4
    ok = np.ones(ii[0].shape, dtype=bool)
    off = np.full(ii[0].shape, plan.tbase_off, dtype=np.int64)
    for d in range(5):
        g = ii[d] * (1 if estr is None else int(estr[d])) + coords[d]
        ok &= (g >= 0) & (g < dims[d])
        off += g.astype(np.int64) * strides[d]
    vals = np.where(ok, xbits[np.where(ok, off // 2, 0)], np.uint16(0))
    lin = ((((ii[4] * box[3] + ii[3]) * box[2] + ii[2]) * box[1] + ii[1]) * box[0] + ii[0]).astype(np.int64) * 2
    if plan.swizzle128:
        lin ^= ((lin >> 7) & 7) << 4
    smem = np.zeros(plan.slab_stride // 2, dtype=np.uint16)
    smem[lin // 2] = vals
    return smem


def umma_operand(smem_bits, start, rows, layout, lbo, sbo):
    """[rows, 16] fp32 operand a K=16 tcgen05.mma reads through a K-major shared-memory descriptor."""
    r = np.arange(rows)[:, None]
    k = np.arange(16)[None, :]
    if layout == 2:
        byte = start + (r >> 3) * sbo + (r & 7) * 128 + k * 2
        byte = byte ^ (((byte >> 7) & 7) << 4)
    else:
        byte = start + (r >> 3) * sbo + (k >> 3) * lbo + (r & 7) * 16 + (k & 7) * 2
    return bits_to_f32(smem_bits[byte // 2])


def stream_block(w_std, k_pad, n0, n_tile, k0):
    """Shared-memory image of one streamed weight block: rows n0..n0+n_tile, K columns k0..k0+64 of the standard
    packed weights, as the 2-D SWIZZLE_128B TMA box {64, n_tile} lays it out."""
    w = w_std.reshape(-1, k_pad)[n0:n0 + n_tile, k0:k0 + 64]
    byte = (np.arange(n_tile)[:, None] * 128 + np.arange(64)[None, :] * 2).astype(np.int64)
    byte ^= ((byte >> 7) & 7) << 4
    slot = np.zeros(n_tile * 64 + 64, dtype=np.uint16)
    slot[byte // 2] = w
    return slot


def simulate_tiles(plan, xbits, image_bits, bias, tiles, k_pad=0, estr=None):
    """Run the plan for the given tile indices; returns {tile: (n[128*tm], tz, oy[128*tm], ox[128*tm], acc[128*tm, n_tile], n0)}
    (oy < 0 or >= OH marks rows the epilogue does not store).
    image_bits: the weight image (resident kinds) or the standard packed weights (streaming kind, with k_pad).
    estr: TMA element strides of a strided 1x1x1 streaming convolution (see tma_box)."""
    out = {}
    nt = plan.n_tile
    img = np.concatenate([image_bits, np.zeros(64, np.uint16)])
    if plan.pair and not plan.b_stream:   # CTA pairs: two per-CTA images of n_tile / 2 rows, read at the same offset
        half = plan.w_bytes // 2
        imgs = [np.concatenate([image_bits[h * half:(h + 1) * half], np.zeros(64, np.uint16)]) for h in (0, 1)]
    for tile in tiles:
        t = tile
        ntile = t % plan.num_n_tiles; t //= plan.num_n_tiles
        n0 = ntile * nt
        tx = t % plan.tiles_x; t //= plan.tiles_x
        ty = t % plan.tiles_y; t //= plan.tiles_y
        tz = t % plan.tiles_z
        n = t // plan.tiles_z
        acc = np.zeros((plan.tm, 128, nt), dtype=np.float64)
        for ks in range(plan.k_stages):
            kt, cb = divmod(ks, plan.cb_n)
            cx, cy = tx * plan.x_step + plan.x_off, ty * plan.y_step + plan.y_off
            cz = tz * plan.z_step + plan.z_off + kt * plan.z_kstep
            coords = (cx * 8, cy, cz, n, 0) if plan.merged_cw else (cb * plan.c_step, cx, cy, cz, n)
            slab = tma_box(xbits, plan, coords, estr)
            tbase = ks * plan.n_grp if plan.tab_per_stage else 0
            for i in range(plan.n_grp * plan.nk):
                grp, kk = divmod(i, plan.nk)
                a_off = plan.tab[2 * (tbase + grp)] + kk * plan.a_kstep
                if plan.b_stream:
                    slot = stream_block(image_bits, k_pad, n0, nt, (kt * plan.n_grp + grp) * plan.cin + cb * 64)
                    B = umma_operand(slot, kk * plan.b_kstep, nt, plan.b_layout, plan.b_lbo, plan.b_sbo).astype(np.float64)
                else:
                    b_off = plan.tab[2 * (tbase + grp) + 1] + kk * plan.b_kstep
                    if plan.pair:
                        B = np.concatenate([umma_operand(im, b_off, nt // 2, plan.b_layout, plan.b_lbo, plan.b_sbo)
                                            for im in imgs]).astype(np.float64)
                    else:
                        B = umma_operand(img, b_off, nt, plan.b_layout, plan.b_lbo, plan.b_sbo).astype(np.float64)
                for h in range(plan.tm):
                    A = umma_operand(slab, a_off + h * plan.half_a_off, 128, plan.a_layout, plan.a_lbo, plan.a_sbo)
                    acc[h] += A.astype(np.float64) @ B.T
        m = np.arange(128)
        g, r = m >> 3, m & 7
        oy = np.concatenate([ty * 16 + g for _ in range(plan.tm)])
        ox = np.concatenate([(tx * plan.tm + h) * 8 + r for h in range(plan.tm)])
        nn = np.full(oy.shape, n, dtype=np.int64)
        if plan.stack_hp:   # stacked rows: the epilogue's (image, row) decode of conv_slab_kernel
            R = oy + plan.stack_ph
            nn = R // plan.stack_hp
            oy = R - nn * plan.stack_hp - plan.stack_ph
            oy = np.where(nn < plan.stack_n, oy, -1)   # rows past the last image are never stored
            nn = np.minimum(nn, plan.stack_n - 1)
        out[tile] = (nn, tz, oy, ox, acc.reshape(plan.tm * 128, nt) + bias[None, n0:n0 + nt].astype(np.float64), n0)
    return out


def simulate_tiles_kx(plan, xbits, image_bits, bias, tiles, cp):
    """The KX kind (TEDSPAD_SLAB_3X3_KX_PAIR): replay the plan for the given tiles and apply slab_epilogue_kx's
    neighbour sums.  image_bits: the two per-CTA halves back to back (each plan.w_bytes); the pair's MMA reads GEMM
    columns [0, N/2) from the leader's image and [N/2, N) from the peer's at the SAME offset.
    Returns {tile: (n[112], oy[112], ox[112], out[112, cp])} for the 8 x 14 output pixels of the tile."""
    assert plan.tm == 1 and plan.n_tile == 3 * cp and plan.pair == 1 and plan.a_sbo == 1024
    half = plan.w_bytes // 2
    imgs = [np.concatenate([image_bits[h * half:(h + 1) * half], np.zeros(64, np.uint16)]) for h in (0, 1)]
    rows = plan.n_tile // 2
    out = {}
    for tile in tiles:
        t = tile
        tx = t % plan.tiles_x; t //= plan.tiles_x
        ty = t % plan.tiles_y; t //= plan.tiles_y
        n = t
        acc = np.zeros((128, plan.n_tile), dtype=np.float64)
        for ks in range(plan.k_stages):
            coords = (ks * plan.c_step, tx * plan.x_step + plan.x_off, ty * plan.y_step + plan.y_off, plan.z_off, n)
            slab = tma_box(xbits, plan, coords)
            for i in range(plan.n_grp * plan.nk):
                grp, kk = divmod(i, plan.nk)
                a_off = plan.tab[2 * (ks * plan.n_grp + grp)] + kk * plan.a_kstep
                b_off = plan.tab[2 * (ks * plan.n_grp + grp) + 1] + kk * plan.b_kstep
                A = umma_operand(slab, a_off, 128, plan.a_layout, plan.a_lbo, plan.a_sbo).astype(np.float64)
                B = np.concatenate([umma_operand(img, b_off, rows, plan.b_layout, plan.b_lbo, plan.b_sbo) for img in imgs])
                acc += A @ B.astype(np.float64).T
        d = acc.reshape(8, 16, 3, cp)                       # [tile row][slab column][filter column][channel]
        res = d[:, 0:14, 0] + d[:, 1:15, 1] + d[:, 2:16, 2] + bias[None, None, :cp].astype(np.float64)
        row, j = np.meshgrid(np.arange(8), np.arange(14), indexing="ij")
        oy = (ty * 8 + row).reshape(-1)
        ox = (tx * 14 + j).reshape(-1)
        nn = np.full(oy.shape, n, dtype=np.int64)
        if plan.stack_hp:
            R = oy + plan.stack_ph
            nn = R // plan.stack_hp
            oy = R - nn * plan.stack_hp - plan.stack_ph
            oy = np.where(nn < plan.stack_n, oy, -1)
            nn = np.minimum(nn, plan.stack_n - 1)
        out[tile] = (nn, oy, ox, res.reshape(112, cp))
    return out


# --------------------------------------------------------------------- staged epilogue stores
def stage_rows(q_rows):
    """conv_slab.cu epi_out, staged form: lane r of an epilogue warp writes the 64 bf16 of its pixel as eight 16-byte
    chunks, chunk k at position k ^ (r & 7) of shared-memory row r.  q_rows: uint16 [32, 64] -> the warp's 4 KB tile."""
    smem = np.zeros(32 * 64, dtype=np.uint16)
    for lane in range(32):
        for k in range(8):
            pos = k ^ (lane & 7)
            smem[lane * 64 + pos * 8: lane * 64 + pos * 8 + 8] = q_rows[lane, k * 8:(k + 1) * 8]
    return smem


def tma_store_box(smem_bits, box=(64, 8, 4)):
    """What a SWIZZLE_128B tiled TMA store reads for box element (c, x, y): the same address rule as tma_box (the
    linear offset of the element with bits 4-6 XORed by bits 7-9).  Returns uint16 [y, x, c]."""
    C, X, Y = box
    c, x, y = np.meshgrid(np.arange(C), np.arange(X), np.arange(Y), indexing="ij")
    lin = (((y * X + x) * C + c) * 2).astype(np.int64)
    lin ^= ((lin >> 7) & 7) << 4
    return smem_bits[lin // 2].transpose(2, 1, 0)

