/*
 * ref_harness_main.cpp -- driver for the reference's own test functions.
 *
 * The reference ships its tests as `bool xxxTest()` functions (tests.h:28-40)
 * and a commented-out main (source.cpp:10-26).  This file is that main: it is
 * linked with the reference's tests.cpp, compiled where it lies, and with ONE
 * implementation of compress()/decompress():
 *   ref_tests_oracle : the CPU oracle            (pins the oracle, runs on CPU)
 *   ref_tests_b200   : libwah_b200.so             (the product, needs a B200)
 *   ref_tests_refgpu : the shim-built reference   (context, needs a GPU)
 *
 * usage: ref_tests_xxx [--big] [name ...]
 * prints one line per test:  RESULT <name> <1|0>
 *
 * TEST INFRASTRUCTURE ONLY.
 */
#include <cstdio>
#include <cstring>
#include <iostream>

#include "tests.h"

struct entry { const char *name; bool (*fn)(); bool big; };

static const entry k_tests[] = {
    { "warpCompressionTest", warpCompressionTest, false },
    { "blockCompressionTest", blockCompressionTest, false },
    { "blockMergeTest", blockMergeTest, false },
    { "blockMergeWithOnesStartsTest", blockMergeWithOnesStartsTest, false },
    { "blockMergeAlternatingTest", blockMergeAlternatingTest, false },
    { "blockMergeFinalLiterals", blockMergeFinalLiterals, false },
    { "blockMergeWanderingLiterals", blockMergeWanderingLiterals, false },
    { "multiBlockTest", multiBlockTest, false },
    { "zerosTest", zerosTest, false },
    { "compressAndDecompressTest", compressAndDecompressTest, true },
    { "randomDataTest", randomDataTest, true },
};

int main(int argc, char **argv)
{
    bool big = false;
    int named = 0;
    for (int i = 1; i < argc; i++) {
        if (strcmp(argv[i], "--big") == 0) big = true;
        else named++;
    }
    int failed = 0;
    for (const entry &t : k_tests) {
        bool run = named == 0 ? (big || !t.big) : false;
        for (int i = 1; i < argc && named; i++)
            if (strcmp(argv[i], t.name) == 0) run = true;
        if (!run) continue;
        bool ok = t.fn();
        std::cout << std::endl << "RESULT " << t.name << " " << (ok ? 1 : 0) << std::endl;
        if (!ok) failed++;
    }
    return failed ? 1 : 0;
}
