"""csrc/ref_expf.h (the restatement of glibc's expf used by every sigmoid/tanh on the device) compiled for the HOST
and checked against libm's expf, plus the KAT derived from the reference's own Sigmoid in tests/golden."""
import os
import subprocess

from conftest import ROOT

SRC = r'''
#include <math.h>
#include <stdio.h>
#include "%s/lightctr_b200/csrc/ref_expf.h"
int main() {
    long bad = 0, n = 0;
    for (uint32_t u = 0; u < 0x42b00000u; u += 499) {
        for (int sg = 0; sg < 2; sg++) {
            uint32_t v = u | (sg ? 0x80000000u : 0);
            float x; memcpy(&x, &v, 4);
            if (!(fabsf(x) < 87.0f)) continue;
            n++;
            if (expf(x) != lctr_ref_expf(x)) bad++;
        }
    }
    printf("%%ld %%ld\n", n, bad);
    return 0;
}
'''


def test_ref_expf_bit_exact_vs_libm(tmp_path):
    c = tmp_path / "t.c"
    c.write_text(SRC % ROOT)
    exe = str(tmp_path / "t")
    subprocess.check_call(["/usr/bin/gcc", "-O2", str(c), "-o", exe, "-lm"])
    n, bad = map(int, subprocess.check_output([exe]).split())
    assert n > 4_000_000 and bad == 0
