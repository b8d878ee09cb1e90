// Chess rules on bitboards -- host+device.
// Semantics (including every deviation from FIDE chess) follow the reference's
// engine/games/chess/src/chess_backend.cpp, cited per function; the representation does not:
// the reference scans a 64-byte mailbox, this keeps four bit planes of 4-bit piece codes and
// derives attack sets from per-square ray tables, emitting moves in the reference's
// scan order (square 0..63, then its per-piece direction order, rays outward) because move
// order is part of the search result (mcts.cpp:57,154 break ties by index).
//
// Square i = r*8+c, r = 0 is rank 8 (chess_backend.cpp:445-457); white moves toward r = 0.
#pragma once
#include <stdint.h>

#include "zc_common.cuh"

namespace zc {
namespace chess {

// piece code: bits 0-2 type (1 P, 2 N, 3 B, 4 R, 5 Q, 6 K), bit 3 = black, 0 = empty
enum { PAWN = 1, KNIGHT = 2, BISHOP = 3, ROOK = 4, QUEEN = 5, KING = 6 };

struct Board {
    uint64_t p0, p1, p2, p3;   // bit planes of the piece code
};
// misc byte carried in the node header: bit 0 turn (0 white), bits 1-4 w_ck, w_cq, b_ck, b_cq
// (state.h:11-13).  The castling flags only feed state_to_tensor; castling is never generated.
constexpr uint32_t MISC_TURN = 1u, MISC_WCK = 2u, MISC_WCQ = 4u, MISC_BCK = 8u, MISC_BCQ = 16u;

constexpr uint64_t FILE_A = 0x0101010101010101ull, FILE_H = 0x8080808080808080ull;

ZC_HD uint64_t bit(int sq) { return 1ull << sq; }
ZC_HD uint64_t occupied(const Board& b) { return b.p0 | b.p1 | b.p2; }
ZC_HD uint64_t of_type(const Board& b, int t) {
    return ((t & 1) ? b.p0 : ~b.p0) & ((t & 2) ? b.p1 : ~b.p1) & ((t & 4) ? b.p2 : ~b.p2);
}
ZC_HD uint64_t of_side(const Board& b, int side) { return occupied(b) & (side ? b.p3 : ~b.p3); }
ZC_HD int piece_at(const Board& b, int sq) {
    return (int)(((b.p0 >> sq) & 1) | (((b.p1 >> sq) & 1) << 1) | (((b.p2 >> sq) & 1) << 2) | (((b.p3 >> sq) & 1) << 3));
}
ZC_HD void put_piece(Board& b, int sq, int code) {
    const uint64_t m = ~bit(sq);
    b.p0 = (b.p0 & m) | ((uint64_t)(code & 1) << sq);
    b.p1 = (b.p1 & m) | ((uint64_t)((code >> 1) & 1) << sq);
    b.p2 = (b.p2 & m) | ((uint64_t)((code >> 2) & 1) << sq);
    b.p3 = (b.p3 & m) | ((uint64_t)((code >> 3) & 1) << sq);
}
ZC_HD int code_of_char(uint8_t ch) {
    switch (ch) {
    case 'P': return 1;  case 'N': return 2;  case 'B': return 3;  case 'R': return 4;  case 'Q': return 5;  case 'K': return 6;
    case 'p': return 9;  case 'n': return 10; case 'b': return 11; case 'r': return 12; case 'q': return 13; case 'k': return 14;
    default: return 0;   // ' ' or 0 (chess_backend.cpp:40-43); any other byte is treated as empty
    }
}
ZC_HD uint8_t char_of_code(int code) {
    const char* t = " PNBRQK  pnbrqk ";
    return (uint8_t)t[code & 15];
}
// captured-piece value attached to a move (chess_backend.cpp:50-64)
ZC_HD int capture_value(int code) {
    switch (code & 7) {
    case PAWN: return 1;  case KNIGHT: case BISHOP: return 3;  case ROOK: return 5;  case QUEEN: return 9;  case KING: return 100;
    default: return 0;
    }
}

// ---- per-square masks (tools/gen_chess_tables.py).  Direction ids follow the reference's queen_dirs
// order (chess_backend.cpp:27-30):  0 (-1,-1)  1 (-1,+1)  2 (+1,-1)  3 (+1,+1)  4 (-1,0)  5 (+1,0)  6 (0,-1)  7 (0,+1)
#include "chess_tables.inc"
#ifdef __CUDACC__
__device__ __align__(64) const uint64_t d_rays_t[64][8] = ZC_RAY_TABLE_T;
__device__ const uint64_t d_knight[64] = ZC_KNIGHT_TABLE;
__device__ const uint64_t d_king[64] = ZC_KING_TABLE;
__device__ __align__(32) const uint64_t d_lines[64][4] = ZC_LINE_TABLE;     // diagonal, anti-diagonal, file, rank through a square
__device__ const uint64_t d_before[64][64] = ZC_BEFORE_TABLE;               // emission-order predecessors (32 KB)
#endif
alignas(64) static const uint64_t h_rays_t[64][8] = ZC_RAY_TABLE_T;
static const uint64_t h_knight[64] = ZC_KNIGHT_TABLE;
static const uint64_t h_king[64] = ZC_KING_TABLE;

// r[d] = squares strictly beyond sq in direction d (d in [D0, D0 + N)): the rays of a square share one
// 64-byte line, fetched with independent 16-byte loads so a caller pays the load latency once
template <int D0, int N>
ZC_HD void rays_of(int sq, uint64_t* r) {
#ifdef __CUDA_ARCH__
    const ulonglong2* line = reinterpret_cast<const ulonglong2*>(&d_rays_t[sq][D0]);
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
        const ulonglong2 v = __ldg(line + i);
        r[2 * i] = v.x;
        r[2 * i + 1] = v.y;
    }
#else
    for (int i = 0; i < N; ++i) r[i] = h_rays_t[sq][D0 + i];
#endif
}
ZC_HD uint64_t knight_targets(int sq) {
#ifdef __CUDA_ARCH__
    return d_knight[sq];
#else
    return h_knight[sq];
#endif
}
ZC_HD uint64_t king_targets(int sq) {
#ifdef __CUDA_ARCH__
    return d_king[sq];
#else
    return h_king[sq];
#endif
}
// directions along which the square index grows
ZC_HD constexpr bool dir_ascending(int d) { return d == 2 || d == 3 || d == 5 || d == 7; }
// the first occupied square met along a ray, as a one-bit mask (0 if the ray is empty): lowest blocker for
// an ascending direction, highest for a descending one
ZC_HD uint64_t first_blocker(int d, uint64_t blockers) {
    if (dir_ascending(d)) return blockers & (0 - blockers);
#ifdef __CUDA_ARCH__
    const uint64_t r = __brevll(blockers);                   // highest bit = lowest bit of the bit-reversed word
    return __brevll(r & (0 - r));
#else
    return blockers ? bit(63 - zc_clz64(blockers)) : 0ull;
#endif
}
// is the first occupied square met along a ray a member of `set` (a subset of the blockers)?  For a descending
// direction that is "the highest bit of blockers belongs to set": set's part outweighs the rest as a number.
ZC_HD bool first_blocker_in(int d, uint64_t blockers, uint64_t set) {
    const uint64_t a = blockers & set;
    if (dir_ascending(d)) return ((blockers & (0 - blockers)) & a) != 0;
    return a > (blockers ^ a);
}
// squares of `ray` (the squares beyond some origin in direction d) up to and including the first blocker
ZC_HD uint64_t ray_until_blocker(int d, uint64_t ray, uint64_t occ) {
    const uint64_t f = first_blocker(d, ray & occ);
    if (dir_ascending(d)) return ray & ((f << 1) - 1);          // f == 0: (0 << 1) - 1 keeps the whole ray
    return f ? ray & ~(f - 1) : ray;
}

#ifdef __CUDACC__
// Slider attacks along one line through a square as two subtractions ("o - 2s" in both bit orders): with o the occupied
// squares of the line (without the slider's own square), s the slider's bit,
//   attacks = ((o - 2s) ^ reverse(reverse(o) - 2 reverse(s))) & line
// reaches up to and including the first blocker on either side.  One expression per LINE instead of one scan per
// direction: a queen's reach is four of these (validated against ray walks for every square and line, tools/ and tests).
struct LineCtx {
    uint64_t line[4];      // d_lines[sq]
    uint64_t s2, s2r;      // 2 * bit(sq), 2 * bit(63 - sq)
};
__device__ __forceinline__ LineCtx line_ctx(int sq) {
    LineCtx c;
    const ulonglong4 v = *reinterpret_cast<const ulonglong4*>(&d_lines[sq][0]);
    c.line[0] = v.x; c.line[1] = v.y; c.line[2] = v.z; c.line[3] = v.w;
    c.s2 = 2ull << sq;
    c.s2r = 2ull << (63 - sq);
    return c;
}
__device__ __forceinline__ uint64_t line_attacks(const LineCtx& c, int l, uint64_t occ) {
    const uint64_t o = occ & c.line[l];
    return ((o - c.s2) ^ __brevll(__brevll(o) - c.s2r)) & c.line[l];
}
#endif

// chess_backend.cpp:85-144 -- is the king of `side` on square ksq attacked, given the enemy sets
// (all subsets of occ)?  Branch-free over the eight rays: the first piece met along a ray attacks iff it
// is an enemy slider of that ray's kind.
ZC_HD_CALL bool square_attacked(int side, int ksq, uint64_t occ, uint64_t e_pawn, uint64_t e_knight, uint64_t e_diag,
                                uint64_t e_orth, uint64_t e_king) {
    const uint64_t k = bit(ksq);
    // pawns: a white king looks one row up (r-1) for 'p', a black king one row down for 'P'
    const uint64_t pawn_from = side == 0 ? (((k >> 9) & ~FILE_H) | ((k >> 7) & ~FILE_A))
                                         : (((k << 7) & ~FILE_H) | ((k << 9) & ~FILE_A));
    uint64_t hit = (pawn_from & e_pawn) | (knight_targets(ksq) & e_knight) | (king_targets(ksq) & e_king);
#ifdef __CUDA_ARCH__
    if (e_diag | e_orth) {
        const LineCtx lc = line_ctx(ksq);
        if (e_diag) hit |= (line_attacks(lc, 0, occ) | line_attacks(lc, 1, occ)) & e_diag;
        if (e_orth) hit |= (line_attacks(lc, 2, occ) | line_attacks(lc, 3, occ)) & e_orth;
    }
#else
    if (e_diag | e_orth) {
        uint64_t r[8];
        rays_of<0, 8>(ksq, r);
        bool ray_hit = false;
#pragma unroll
        for (int d = 0; d < 8; ++d) ray_hit |= first_blocker_in(d, r[d] & occ, d < 4 ? e_diag : e_orth);
        if (ray_hit) return true;
    }
#endif
    return hit != 0;
}

// The same test for the side to move's king in the current position, plus the own pieces PINNED to it: the
// first piece met along a ray is own and the second is an enemy slider of that ray's kind.  Only such a
// piece can expose the king by leaving its square (legality filter, chess_backend.cpp:345-358).
ZC_HD bool king_danger(int side, int ksq, uint64_t occ, uint64_t own, uint64_t e_pawn, uint64_t e_knight, uint64_t e_diag,
                       uint64_t e_orth, uint64_t e_king, uint64_t& pinned) {
    const uint64_t k = bit(ksq);
    const uint64_t pawn_from = side == 0 ? (((k >> 9) & ~FILE_H) | ((k >> 7) & ~FILE_A))
                                         : (((k << 7) & ~FILE_H) | ((k << 9) & ~FILE_A));
    uint64_t hit = (pawn_from & e_pawn) | (knight_targets(ksq) & e_knight) | (king_targets(ksq) & e_king);
    uint64_t pins = 0;
    if (e_diag | e_orth) {
        uint64_t r[8];
        rays_of<0, 8>(ksq, r);
#pragma unroll
        for (int d = 0; d < 8; ++d) {
            const uint64_t sliders = d < 4 ? e_diag : e_orth;
            if (!(r[d] & sliders)) continue;                     // no such slider on this ray at all
            const uint64_t b = r[d] & occ, f1 = first_blocker(d, b);
            hit |= f1 & sliders;
            if (first_blocker(d, b & ~f1) & sliders) pins |= f1 & own;
        }
    }
    pinned = pins;
    return hit != 0;
}

struct Sets {   // derived once per position
    uint64_t occ, own, enemy, e_pawn, e_knight, e_bishop, e_rook, e_queen, e_king, own_king;
};
ZC_HD Sets derive(const Board& b, int turn) {
    Sets s;
    s.occ = occupied(b);
    s.own = of_side(b, turn);
    s.enemy = s.occ & ~s.own;
    s.e_pawn = of_type(b, PAWN) & s.enemy;
    s.e_knight = of_type(b, KNIGHT) & s.enemy;
    s.e_bishop = of_type(b, BISHOP) & s.enemy;
    s.e_rook = of_type(b, ROOK) & s.enemy;
    s.e_queen = of_type(b, QUEEN) & s.enemy;
    s.e_king = of_type(b, KING) & s.enemy;
    s.own_king = of_type(b, KING) & s.own;
    return s;
}

// A side WITHOUT a king (only a hand-made position): find_king leaves (kr, kc) = (-1, -1) (chess_backend.cpp:68-81) and
// king_attacked, whose every read is bounds-checked (:85-144), tests that phantom square: what can "attack" it from the board
// is a pawn on square 0 when black is to move ((kr+1, kc+1), :97), a knight on (0,1) or (1,0), a bishop or queen first
// on the diagonal (0,0), (1,1), ..., or the enemy king on square 0.  Reproduced as is.
// (out of line: a kingless side never occurs in play, and the hot paths must not carry this code)
constexpr uint64_t MAIN_DIAGONAL = 0x8040201008040201ull;
ZC_HD_CALL bool phantom_king_attacked(int side, uint64_t occ, uint64_t e_pawn, uint64_t e_knight, uint64_t e_diag, uint64_t e_king) {
    const uint64_t on_diag = occ & MAIN_DIAGONAL;
    return (((side == 1 ? e_pawn : 0ull) | e_king) & 1ull) != 0 || (e_knight & (bit(1) | bit(8))) != 0 ||
           ((on_diag & (0 - on_diag)) & e_diag) != 0;
}

// the make-move test of a kingless side: the phantom square after the move
ZC_HD_CALL bool kingless_move_ok(uint64_t occ, int turn, int from, int to, uint64_t e_pawn, uint64_t e_knight, uint64_t e_diag, uint64_t e_king) {
    const uint64_t keep = ~bit(to);
    return !phantom_king_attacked(turn, (occ & ~bit(from)) | bit(to), e_pawn & keep, e_knight & keep, e_diag & keep, e_king & keep);
}

// chess_backend.cpp:345-358 -- make the move, find the mover's (first) king, test it.  Precondition: the mover HAS a king
// (callers route a kingless side to kingless_move_ok).
ZC_HD bool move_keeps_king_safe(const Sets& s, int turn, int from, int to, bool king_moves) {
    const uint64_t f = bit(from), t = bit(to);
    const uint64_t kings = king_moves ? ((s.own_king & ~f) | t) : s.own_king;
    const uint64_t occ = (s.occ & ~f) | t, keep = ~t;
    const int ksq = zc_ctz64(kings);
    return !square_attacked(turn, ksq, occ, s.e_pawn & keep, s.e_knight & keep, (s.e_bishop | s.e_queen) & keep,
                            (s.e_rook | s.e_queen) & keep, s.e_king & keep);
}
ZC_HD bool in_check(const Board& b, int turn) {   // king of the side to move attacked now?
    const Sets s = derive(b, turn);
    if (!s.own_king) {                            // the phantom square, inline here (a dozen instructions, no call in a hot function)
        const uint64_t on_diag = s.occ & MAIN_DIAGONAL;
        return ((((turn == 1 ? s.e_pawn : 0ull) | s.e_king) & 1ull) | (s.e_knight & (bit(1) | bit(8))) |
                ((on_diag & (0 - on_diag)) & (s.e_bishop | s.e_queen))) != 0;
    }
    return square_attacked(turn, zc_ctz64(s.own_king), s.occ, s.e_pawn, s.e_knight, s.e_bishop | s.e_queen,
                           s.e_rook | s.e_queen, s.e_king);
}

// chess_backend.cpp:188-198 -- no pawn/rook/queen and at most one minor piece: no moves at all
ZC_HD bool insufficient_material(const Board& b) {
    const uint64_t heavy = of_type(b, PAWN) | of_type(b, ROOK) | of_type(b, QUEEN);
    const uint64_t minor = of_type(b, KNIGHT) | of_type(b, BISHOP);
    return heavy == 0 && zc_popc64(minor) <= 1;
}

ZC_HD uint16_t pack_move(int from, int to) { return (uint16_t)(from | (to << 6)); }
ZC_HD int move_from(uint16_t m) { return m & 63; }
ZC_HD int move_to(uint16_t m) { return (m >> 6) & 63; }

// chess_backend.cpp:184-360.  Writes the legal moves, in the reference's order, to out[] (room for
// MAX_PSEUDO entries, never null) and returns their number (<= 218).
//
// Pass 1 emits the pseudo-legal moves in the reference's scan order (:203-342).  Pass 2 is the
// reference's stable legality filter (:345-358) with ONE call site of the attack test and an exact
// shortcut: if the mover is not in check, a non-king piece can expose its king by leaving its square only
// if it is pinned (alone between the king and an enemy slider on a line the slider moves along); its
// destination can only block lines and a capture only removes an attacker.  So unpinned pieces' moves
// are legal without testing; king moves, pinned pieces and every move while in check take the full test.
constexpr int MAX_PSEUDO = 256;
constexpr uint16_t MOVE_KING_FLAG = 1u << 12;

// `stride`: distance between consecutive entries of out[] (32 when the lanes of a warp interleave their lists so
// that lane-parallel accesses to entry i coalesce)
ZC_HD int generate(const Board& b, int turn, uint16_t* out, int stride = 1) {
    if (insufficient_material(b)) return 0;
    const Sets s = derive(b, turn);
    const uint64_t empty = ~s.occ;
    const uint64_t targets_ok = ~s.own & ~s.e_king;       // a king is never captured (:240,261,306,331)
    int n = 0;
    uint64_t movers = s.own;
    while (movers) {
        const int sq = zc_ctz64(movers);
        movers &= movers - 1;
        const int type = piece_at(b, sq) & 7, r = sq >> 3, c = sq & 7;
        if (type == PAWN) {                                             // :213-252
            const int dir = turn == 0 ? -1 : 1, home = turn == 0 ? 6 : 1;
            const int nr = r + dir;
            if (nr >= 0 && nr < 8) {
                const int one = nr * 8 + c;
                if (empty >> one & 1) {
                    out[(n++) * stride] = pack_move(sq, one);
                    const int two = one + dir * 8;
                    if (r == home && (empty >> two & 1)) out[(n++) * stride] = pack_move(sq, two);
                }
                const uint64_t capturable = s.enemy & ~s.e_king;
                if (c > 0 && (capturable >> (one - 1) & 1)) out[(n++) * stride] = pack_move(sq, one - 1);
                if (c < 7 && (capturable >> (one + 1) & 1)) out[(n++) * stride] = pack_move(sq, one + 1);
            }
        } else if (type == KNIGHT) {                                    // :255-275; knight_dirs order == ascending target
            uint64_t tg = knight_targets(sq) & targets_ok;
            while (tg) {
                const int t = zc_ctz64(tg);
                tg &= tg - 1;
                out[(n++) * stride] = pack_move(sq, t);
            }
        } else if (type == KING) {                                      // :322-340, king_dirs order (:31-34)
            for (int d = 0; d < 8; ++d) {
                const int dr = d < 4 ? (d < 2 ? -1 : 1) : (d == 4 ? -1 : d == 5 ? 1 : 0);
                const int dc = d < 4 ? ((d & 1) ? 1 : -1) : (d == 6 ? -1 : d == 7 ? 1 : 0);
                const int rr = r + dr, cc = c + dc;
                if (rr < 0 || rr > 7 || cc < 0 || cc > 7) continue;
                const int t = rr * 8 + cc;
                if (targets_ok >> t & 1) out[(n++) * stride] = (uint16_t)(pack_move(sq, t) | MOVE_KING_FLAG);
            }
        } else if (type == BISHOP || type == ROOK || type == QUEEN) {   // :278-319
            uint64_t rr[8];
            rays_of<0, 8>(sq, rr);
            const int d0 = type == ROOK ? 4 : 0, d1 = type == BISHOP ? 4 : 8;
#pragma unroll
            for (int d = 0; d < 8; ++d) {
                if (d < d0 || d >= d1) continue;
                uint64_t tg = ray_until_blocker(d, rr[d], s.occ) & targets_ok;
                while (tg) {                                            // outward from the piece
                    const int t = dir_ascending(d) ? zc_ctz64(tg) : 63 - zc_clz64(tg);
                    tg &= ~bit(t);
                    out[(n++) * stride] = pack_move(sq, t);
                }
            }
        }
    }
    // ---- pass 2: legality filter, stable, in place
    if (!s.own_king) {                       // a side without a king: every move takes the make-move test against the phantom square
        int m = 0;
        for (int i = 0; i < n; ++i) {
            const uint16_t mv = out[i * stride];
            if (kingless_move_ok(s.occ, turn, mv & 63, (mv >> 6) & 63, s.e_pawn, s.e_knight, s.e_bishop | s.e_queen, s.e_king))
                out[(m++) * stride] = (uint16_t)(mv & 0x0FFF);
        }
        return m;
    }
    const int ksq = zc_ctz64(s.own_king);    // find_king: the first king in index order (:68-81)
    uint64_t pinned;
    const bool checked = king_danger(turn, ksq, s.occ, s.own, s.e_pawn, s.e_knight, s.e_bishop | s.e_queen,
                                     s.e_rook | s.e_queen, s.e_king, pinned);
    int m = 0;
    for (int i = 0; i < n; ++i) {
        const uint16_t mv = out[i * stride];
        const int from = mv & 63, to = (mv >> 6) & 63;
        const bool king_moves = (mv & MOVE_KING_FLAG) != 0;
        bool ok = true;
        if (checked || king_moves || (pinned >> from & 1)) ok = move_keeps_king_safe(s, turn, from, to, king_moves);
        if (ok) out[(m++) * stride] = (uint16_t)(mv & 0x0FFF);
    }
    return m;
}

#ifdef __CUDACC__
// out of line: reached only for a contrived position with more than 32 pieces of one colour
__device__ __noinline__ int generate_cold(const Board& b, int turn, uint16_t* out, int stride) {
    if (out == nullptr) zc_layout_pad<ZC_PAD_COLD>();      // never true (zc_common.cuh: code layout); this function sits between the hot ones
    return generate(b, turn, out, stride);
}

// The king of the side to move: is it attacked, and which own pieces are PINNED to it (the first piece met from the king
// along a line is own and the next one beyond it is an enemy slider of that line's kind)?  Line form of king_danger().
__device__ __forceinline__ bool king_danger_lines(int side, int ksq, uint64_t occ, uint64_t own, uint64_t e_pawn, uint64_t e_knight,
                                                  uint64_t e_diag, uint64_t e_orth, uint64_t e_king, uint64_t& pinned) {
    const uint64_t k = bit(ksq);
    const uint64_t pawn_from = side == 0 ? (((k >> 9) & ~FILE_H) | ((k >> 7) & ~FILE_A))
                                         : (((k << 7) & ~FILE_H) | ((k << 9) & ~FILE_A));
    uint64_t hit = (pawn_from & e_pawn) | (knight_targets(ksq) & e_knight) | (king_targets(ksq) & e_king);
    uint64_t pins = 0;
    if (e_diag | e_orth) {
        const LineCtx lc = line_ctx(ksq);
        const uint64_t above = ~((k << 1) - 1);              // squares with a higher index than the king: one side of every line
#pragma unroll 1
        for (int l = 0; l < 4; ++l) {
            const uint64_t sliders = l < 2 ? e_diag : e_orth;
            if (!(lc.line[l] & sliders)) continue;           // no such slider on this line at all
            const uint64_t att = line_attacks(lc, l, occ);
            hit |= att & sliders;
            const uint64_t blockers = att & own;             // own pieces in direct view of the king (at most one per side)
            if (!blockers) continue;
            const uint64_t pinners = line_attacks(lc, l, occ ^ blockers) & ~att & sliders;   // seen only through them
            if (pinners & above) pins |= blockers & above;
            if (pinners & ~above) pins |= blockers & ~above;
        }
    }
    pinned = pins;
    return hit != 0;
}

// index of the n-th (0-based) set bit of a 64-bit word, lowest first; n < popcount(v).  Binary search over popcounts
// (the __fns intrinsic is a long software sequence).
__device__ __forceinline__ int nth_set_bit(uint64_t v, int n) {
    uint32_t w = (uint32_t)v;
    int base = 0;
    int c = __popc(w);
    if (n >= c) { n -= c; w = (uint32_t)(v >> 32); base = 32; }
#pragma unroll
    for (int width = 16; width >= 1; width >>= 1) {
        const uint32_t low = w & ((1u << width) - 1u);
        c = __popc(low);
        if (n >= c) { n -= c; w >>= width; base += width; }
        else w = low;
    }
    return base;
}

// The same list as generate(), produced by a whole warp for ONE position: lane r owns the side to move's r-th piece in
// square order (the order the reference scans, chess_backend.cpp:203) and holds ALL its targets as one bitboard -- slider
// reach from four line expressions, leapers from tables, the (up to four) pawn moves by hand.  The targets are filtered
// with the same legality rule as generate(); an exclusive scan of the per-piece counts places every piece's moves, and a
// move's position among its piece's moves -- the reference's per-piece direction order, rays outward -- is the number
// of the piece's targets that precede it: popcount(targets & d_before[sq][t]) (knights: ascending squares; pawns: push,
// double push, capture left, capture right).  All lanes must call it with the same arguments; out[i * stride] = move i.
//
// any_only (warp-uniform; the caller knows the side to move is IN CHECK): answer only "is there a legal move?" -- 1 or 0,
// nothing written to out[].  check_win (chess_backend.cpp:404-412) needs no more than that for a leaf, and in check every
// move takes the full make-move test, so stopping at the first legal one found by any lane skips most of the work.
__device__ __forceinline__ int generate_warp(const Board& b, int turn, uint16_t* out, int stride, int lane, bool any_only = false) {
    if (insufficient_material(b)) return 0;
    const Sets s = derive(b, turn);
    const int n_own = zc_popc64(s.own);
    if (n_own > 32 || !s.own_king) {                   // more pieces than lanes, or no king (only a contrived FEN): one lane does it
        int n = 0;
        if (out == nullptr) zc_layout_pad<ZC_PAD_GEN>();   // never true (zc_common.cuh: code layout)
        if (lane == 0) n = generate_cold(b, turn, out, stride);
        return __shfl_sync(0xFFFFFFFFu, n, 0);
    }
    const uint64_t empty = ~s.occ;
    const uint64_t targets_ok = ~s.own & ~s.e_king;    // a king is never captured (:240,261,306,331)
    const bool has_king = s.own_king != 0;
    const int ksq = has_king ? zc_ctz64(s.own_king) : 0;   // find_king: the first king in index order (:68-81)
    uint64_t pinned = 0;
    const bool checked = king_danger_lines(turn, ksq, s.occ, s.own, s.e_pawn, s.e_knight, s.e_bishop | s.e_queen,
                                           s.e_rook | s.e_queen, s.e_king, pinned);
    uint64_t targets = 0;                              // this lane's piece may move to these squares (before the legality filter)
    int sq = 0, type = 0;
    uint32_t pawn_order = 0;                           // 4 x 6 bits: the pawn's targets in emission order (63 = none)
    const bool mine = lane < n_own;
    if (mine) {
        sq = nth_set_bit(s.own, lane);
        type = piece_at(b, sq) & 7;
        if (type == PAWN) {                            // single push, double push, capture dc=-1, capture dc=+1 (:213-252)
            const int r = sq >> 3, c = sq & 7;
            const int dir = turn == 0 ? -1 : 1, home = turn == 0 ? 6 : 1;
            const int nr = r + dir;
            if (nr >= 0 && nr < 8) {
                const int one = nr * 8 + c, two = one + dir * 8;
                const uint64_t capturable = s.enemy & ~s.e_king;
                const bool p1 = (empty >> one & 1) != 0, p2 = p1 && r == home && (empty >> two & 1) != 0;
                const bool cl = c > 0 && (capturable >> (one - 1) & 1) != 0, cr = c < 7 && (capturable >> (one + 1) & 1) != 0;
                targets = (p1 ? bit(one) : 0) | (p2 ? bit(two) : 0) | (cl ? bit(one - 1) : 0) | (cr ? bit(one + 1) : 0);
                pawn_order = (uint32_t)one | ((uint32_t)(two & 63) << 6) | ((uint32_t)((one - 1) & 63) << 12) | ((uint32_t)((one + 1) & 63) << 18);
            }
        } else if (type == KNIGHT) {
            targets = knight_targets(sq) & targets_ok;
        } else if (type == KING) {
            targets = king_targets(sq) & targets_ok;
        } else if (type != 0) {                        // bishop, rook, queen
            const LineCtx lc = line_ctx(sq);
            if (type != ROOK) targets |= line_attacks(lc, 0, s.occ) | line_attacks(lc, 1, s.occ);
            if (type != BISHOP) targets |= line_attacks(lc, 2, s.occ) | line_attacks(lc, 3, s.occ);
            targets &= targets_ok;
        }
    }
    // The reference's make-move test (:345-358), ONE call site fed by a per-lane work list.  The common case --
    // not in check, so only the king's own steps (and a rare pinned piece) need it -- hands the tested king's
    // candidate steps to lanes 24..31, one each, which own no piece when the side has at most 24 of them; in every
    // other case a lane walks the targets of its own piece.
    const bool spread_king = has_king && !checked && n_own <= 24;
    const bool own_king_lane = mine && type == KING && sq == ksq;
    uint64_t work = 0;                                 // targets this lane still has to test
    int wfrom = sq;
    bool wking = type == KING;
    {
        if (spread_king) {
            const uint64_t ktg = king_targets(ksq) & targets_ok;
            if (lane >= 24) {
                if (lane - 24 < zc_popc64(ktg)) work = bit(nth_set_bit(ktg, lane - 24));
                wfrom = ksq;
                wking = true;
            } else if (mine && !own_king_lane && (type == KING || (pinned >> sq & 1))) {
                work = targets;
            }
        } else if (mine && (checked || type == KING || (pinned >> sq & 1))) {
            work = targets;
        }
    }
    uint64_t bad = 0;
    bool found = false;
    while (__any_sync(0xFFFFFFFFu, work != 0)) {
        if (work) {
            const int t = zc_ctz64(work);
            work &= work - 1;
            if (!move_keeps_king_safe(s, turn, wfrom, t, wking)) bad |= bit(t);
            else found = true;
        }
        if (any_only && __any_sync(0xFFFFFFFFu, found)) return 1;
    }
    if (any_only) return 0;
    if (spread_king) {                                 // lanes 24..31 report their step to the king's lane
        const bool rep = lane >= 24;
        const uint32_t bad_lo = __reduce_or_sync(0xFFFFFFFFu, rep ? (uint32_t)bad : 0u);
        const uint32_t bad_hi = __reduce_or_sync(0xFFFFFFFFu, rep ? (uint32_t)(bad >> 32) : 0u);
        if (own_king_lane) targets &= ~(((uint64_t)bad_hi << 32) | bad_lo);
        if (rep) targets = 0;
    }
    if (mine && !(spread_king && lane >= 24)) targets &= ~bad;
    const int cnt = zc_popc64(targets);
    int x = cnt;                                       // exclusive scan over lanes = over pieces in square order
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int y = __shfl_up_sync(0xFFFFFFFFu, x, d);
        if (lane >= d) x += y;
    }
    const int total = __shfl_sync(0xFFFFFFFFu, x, 31);
    const int pos = x - cnt;
    // Emission, any order (every move knows its own index).  With at most 16 pieces a side -- always, in a game -- lanes
    // 16..31 own no piece: lane L + 16 takes over the upper half of lane L's target squares, so the walk over the set bits is
    // half as long and 32 bits wide.
    const bool split = n_own <= 16;
    const int src = split ? (lane & 15) : lane;
    const uint64_t e_targets = __shfl_sync(0xFFFFFFFFu, targets, src);
    const int e_sq = __shfl_sync(0xFFFFFFFFu, sq, src), e_type = __shfl_sync(0xFFFFFFFFu, type, src), e_pos = __shfl_sync(0xFFFFFFFFu, pos, src);
    const uint32_t e_pawn = __shfl_sync(0xFFFFFFFFu, pawn_order, src);
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
        if (split && half != (lane >> 4)) continue;
        uint32_t tg = half ? (uint32_t)(e_targets >> 32) : (uint32_t)e_targets;
        while (tg) {
            const int t = (__ffs((int)tg) - 1) + 32 * half;
            tg &= tg - 1;
            int rank;
            if (e_type == PAWN) {
                rank = 0;
#pragma unroll
                for (int q = 0; q < 3; ++q) {          // targets listed before t in the pawn's fixed order
                    const int tq = (int)(e_pawn >> (6 * q)) & 63;
                    if (tq == t) break;
                    rank += (int)(e_targets >> tq & 1);
                }
            } else {
                const uint64_t before = e_type == KNIGHT ? bit(t) - 1 : d_before[e_sq][t];
                rank = zc_popc64(e_targets & before);
            }
            out[(size_t)(e_pos + rank) * stride] = pack_move(e_sq, t);
        }
    }
    __syncwarp();
    return total;
}
#endif

// chess_backend.cpp:364-400 on the board planes; flags in/out through `misc`.
ZC_HD Board play(const Board& b, uint32_t misc, int from, int to, uint32_t& misc_out) {
    Board n = b;
    const int pc = piece_at(b, from), fc = from & 7, tc = to & 7, tr = to >> 3;
    uint32_t m = misc ^ MISC_TURN;
    if (pc == KING || (pc == ROOK && fc == 7)) m &= ~MISC_WCK;               // :382-385 (any row, as written)
    if (pc == KING || (pc == ROOK && fc == 0)) m &= ~MISC_WCQ;
    if (pc == (KING | 8) || (pc == (ROOK | 8) && fc == 7)) m &= ~MISC_BCK;
    if (pc == (KING | 8) || (pc == (ROOK | 8) && fc == 0)) m &= ~MISC_BCQ;
    // castling rook hop, only reachable when a two-file king move is fed in (:388-391)
    if (pc == KING && tc - fc == 2) { put_piece(n, 61, ROOK); put_piece(n, 63, 0); }
    if (pc == (KING | 8) && tc - fc == 2) { put_piece(n, 5, ROOK | 8); put_piece(n, 7, 0); }
    if (pc == KING && tc - fc == -2) { put_piece(n, 59, ROOK); put_piece(n, 56, 0); }
    if (pc == (KING | 8) && tc - fc == -2) { put_piece(n, 3, ROOK | 8); put_piece(n, 0, 0); }
    put_piece(n, to, pc);                                                     // :393-394, in the reference's order: a move
    put_piece(n, from, 0);                                                    // onto its own square empties the square ...
    if (tr == 0 && pc == PAWN) put_piece(n, to, QUEEN);                       // :396-397 ... unless it "promotes" there
    if (tr == 7 && pc == (PAWN | 8)) put_piece(n, to, QUEEN | 8);
    misc_out = m;
    return n;
}
// The same for a move the generator produced.  Generated king moves are single steps (castling is never generated,
// :322-340), so the rook hop of :388-391 cannot trigger and the move is: lift the piece's four code bits off `from`, drop
// them on `to` (replacing what stood there), promote a pawn that reaches the last row (:396-397: pawn 001 -> queen 101).
ZC_HD Board play_generated(const Board& b, uint32_t misc, int from, int to, uint32_t& misc_out) {
    const uint64_t f = bit(from), t = bit(to), keep = ~(f | t);
    const int pc = piece_at(b, from), fc = from & 7, tr = to >> 3;
    uint32_t m = misc ^ MISC_TURN;
    if (pc == KING || (pc == ROOK && fc == 7)) m &= ~MISC_WCK;               // :382-385 (any row, as written)
    if (pc == KING || (pc == ROOK && fc == 0)) m &= ~MISC_WCQ;
    if (pc == (KING | 8) || (pc == (ROOK | 8) && fc == 7)) m &= ~MISC_BCK;
    if (pc == (KING | 8) || (pc == (ROOK | 8) && fc == 0)) m &= ~MISC_BCQ;
    Board n;
    n.p0 = (b.p0 & keep) | ((pc & 1) ? t : 0);
    n.p1 = (b.p1 & keep) | ((pc & 2) ? t : 0);
    n.p2 = (b.p2 & keep) | ((pc & 4) ? t : 0);
    n.p3 = (b.p3 & keep) | ((pc & 8) ? t : 0);
    if ((tr == 0 && pc == PAWN) || (tr == 7 && pc == (PAWN | 8))) n.p2 |= t;
    misc_out = m;
    return n;
}
// does this move reset the fifty-ply counter?  (:380: pawn move or any capture)
ZC_HD bool resets_fifty(const Board& b, int from, int to) {
    return (piece_at(b, from) & 7) == PAWN || piece_at(b, to) != 0;
}

// sum of piece values, white positive (value_functions.py:52-54)
ZC_HD int material(const Board& b) {
    const uint64_t w = ~b.p3, k = b.p3;
    const uint64_t P = of_type(b, PAWN), Nn = of_type(b, KNIGHT) | of_type(b, BISHOP), R = of_type(b, ROOK), Q = of_type(b, QUEEN);
    return (zc_popc64(P & w) - zc_popc64(P & k)) + 3 * (zc_popc64(Nn & w) - zc_popc64(Nn & k)) +
           5 * (zc_popc64(R & w) - zc_popc64(R & k)) + 9 * (zc_popc64(Q & w) - zc_popc64(Q & k));
}
// crude_chess_score (value_functions.py:49-55) given the number of legal moves of the state:
// check_win = no moves and king attacked (chess_backend.cpp:404-412) -> +1000 (sic, for the MATED side)
ZC_HD double crude_score(const Board& b, int turn, int n_moves) {
    if (n_moves == 0 && in_check(b, turn)) return 1000.0;
    return (double)((turn ? -1 : 1) * material(b));
}

}  // namespace chess
}  // namespace zc
