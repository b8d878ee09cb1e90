"""Host check of chess::play_generated (the lean move application the search uses for a node's own generated moves) against
chess::play (the full chess_backend.cpp:364-400 semantics incl. the castling rook hop): identical boards and flags for every
generated move of random games.  The rule code is __host__ __device__, so g++ compiles the very functions the kernels run."""
import os
import subprocess
import sys
import textwrap

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = textwrap.dedent('''
    #include <cstdio>
    #include "zeroclone_b200/csrc/chess_rules.cuh"
    using namespace zc::chess;
    int main() {
        unsigned long long seed = 12345, bad = 0, n = 0;
        auto rnd = [&]() { seed = seed * 6364136223846793005ull + 1442695040888963407ull; return (unsigned)(seed >> 33); };
        for (int game = 0; game < 400; ++game) {
            Board b = {0, 0, 0, 0};
            const char* init = "rnbqkbnrpppppppp                                PPPPPPPPRNBQKBNR";
            for (int i = 0; i < 64; ++i) put_piece(b, i, code_of_char((unsigned char)init[i]));
            uint32_t misc = MISC_WCK | MISC_WCQ | MISC_BCK | MISC_BCQ;
            for (int ply = 0; ply < 150; ++ply) {
                uint16_t mv[MAX_PSEUDO];
                const int k = generate(b, (int)(misc & 1u), mv);
                if (k == 0) break;
                for (int i = 0; i < k; ++i) {
                    uint32_t m1, m2;
                    const Board a = play(b, misc, move_from(mv[i]), move_to(mv[i]), m1);
                    const Board c = play_generated(b, misc, move_from(mv[i]), move_to(mv[i]), m2);
                    ++n;
                    if (a.p0 != c.p0 || a.p1 != c.p1 || a.p2 != c.p2 || a.p3 != c.p3 || m1 != m2) ++bad;
                }
                uint32_t nm;
                const int pick = (int)(rnd() % (unsigned)k);
                b = play(b, misc, move_from(mv[pick]), move_to(mv[pick]), nm);
                misc = nm;
            }
        }
        std::printf("%llu %llu\\n", n, bad);
        return bad != 0;
    }
''')


def test_play_generated_equals_play_on_generated_moves(tmp_path):
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    src, exe = tmp_path / "pg.cpp", tmp_path / "pg"
    src.write_text(SRC)
    build = subprocess.run([cxx, "-O2", "-std=c++17", f"-I{REPO}", "-o", str(exe), str(src)], capture_output=True, text=True)
    if build.returncode != 0:
        pytest.fail(build.stderr[-2000:])
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    n, bad = (int(v) for v in run.stdout.split())
    assert run.returncode == 0 and bad == 0 and n > 500_000, (n, bad)
