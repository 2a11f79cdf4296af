/* `ola prove -i trace.json -o proof.bin` (client/src/main.rs:172-207) as a C host of libola_gpu.so: read the JSON text `ola run`
 * wrote, parse it (ola_trace_from_json), generate the twelve tables on the GPU and prove them there (ola_prove_trace), write
 * Buffer::write_all_proof's bytes; then `ola verify` (main.rs:208-243) on the file just written (ola_verify, host code).
 *   gcc -O2 -I include tests/c/ola_prove_file.c -o ola_prove_file -L olavm_b200 -lola_gpu -Wl,-rpath,$PWD/olavm_b200
 *   ./ola_prove_file trace.json proof.bin
 * Exit code 0 = proved and verified; 2 = the text is not a Trace; 3 = no GPU (there is no CPU fallback). */
#include <stdio.h>
#include <stdlib.h>

#include "ola_gpu.h"

static char* read_file(const char* path, size_t* len) {
    FILE* f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    char* buf = (char*)malloc((size_t)n + 1);
    if (buf && fread(buf, 1, (size_t)n, f) != (size_t)n) {
        free(buf);
        buf = NULL;
    }
    fclose(f);
    *len = (size_t)n;
    return buf;
}

int main(int argc, char** argv) {
    if (argc != 3) {
        fprintf(stderr, "usage: %s trace.json proof.bin\n", argv[0]);
        return 1;
    }
    size_t len = 0;
    char* text = read_file(argv[1], &len);
    if (!text) {
        fprintf(stderr, "cannot read %s\n", argv[1]);
        return 1;
    }
    printf("Input trace file path: %s\n", argv[1]);
    char err[256];
    ola_trace* trace = NULL;
    if (ola_trace_from_json(text, len, &trace, err, sizeof err) != OLA_OK) {
        fprintf(stderr, "not a Trace: %s\n", err);
        return 2;
    }
    free(text);
    for (int t = 0; t < 12; ++t) printf("table %2d: 2^%d rows x %d columns\n", t, ola_trace_table_log_rows(trace, t), ola_table_columns(t));
    ola_ctx* ctx = NULL;
    if (ola_gpu_init(0, &ctx) != OLA_OK) {
        fprintf(stderr, "ola_gpu_init failed: %s\n", ola_gpu_last_error(NULL));
        ola_trace_free(trace);
        return 3;
    }
    size_t cap = (size_t)1 << 24, n = 0;
    uint8_t* proof = (uint8_t*)malloc(cap);
    int rc = ola_prove_trace(ctx, trace, proof, cap, &n);
    ola_trace_free(trace);
    if (rc != OLA_OK) {
        fprintf(stderr, "ola_prove_trace failed (%d): %s\n", rc, ola_gpu_last_error(ctx));
        ola_gpu_destroy(ctx);
        return 4;
    }
    ola_gpu_destroy(ctx);
    FILE* out = fopen(argv[2], "wb");
    if (!out || fwrite(proof, 1, n, out) != n) {
        fprintf(stderr, "cannot write %s\n", argv[2]);
        return 1;
    }
    fclose(out);
    printf("Proof size: %zu bytes\nProve done!\n", n);
    const int ids[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
    rc = ola_verify(ids, 12, proof, n, err, sizeof err);
    free(proof);
    if (rc != OLA_OK) {
        fprintf(stderr, "Verify failed: %s\n", err);
        return 5;
    }
    printf("Verify succeed!\n");
    return 0;
}
