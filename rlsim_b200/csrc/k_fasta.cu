// FASTA record formatting on the device.
//
// Restates Frag.String / Frag.GetSeq (frag.go:71-89,123-126), RevCompDNA (utils.go:37-59) and the
// `fmt.Printf("%s\n", frag.String())` of sampler.go:206:
//     >Fg_<id>_<name> (Strand <s> Offset <start> -- <end>)\n<sequence>\n
// "+" prints plus[start:end]; "-" prints minus[len-end : len-start], i.e. the reverse complement of the same
// interval (non-ACGT letters become 'N').  One warp per record.
#include "kernels.cuh"

namespace rl {

__device__ __forceinline__ uint32_t dec_digits(uint64_t v) {
    uint32_t d = 1;
    while (v >= 10) { v /= 10; d++; }
    return d;
}
__device__ __forceinline__ uint8_t *put_dec(uint8_t *p, uint64_t v) {
    uint32_t d = dec_digits(v);
    for (uint32_t i = 0; i < d; i++) { p[d - 1 - i] = (uint8_t)('0' + v % 10); v /= 10; }
    return p + d;
}
__device__ __forceinline__ uint8_t *put_str(uint8_t *p, const char *s) {
    while (*s) *p++ = (uint8_t)*s++;
    return p;
}

__global__ void k_fasta_sizes(DevSampled s, uint64_t n, const uint64_t *__restrict__ name_off, uint32_t first_index,
                              uint64_t *__restrict__ sizes) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t t = s.tr[i] - first_index;
    uint64_t nl = name_off[t + 1] - name_off[t];
    // ">Fg_" id "_" name " (Strand " c " Offset " start " -- " end ")\n" seq "\n"
    sizes[i] = 4 + dec_digits(s.id[i]) + 1 + nl + 9 + 1 + 8 + dec_digits(s.start[i]) + 4 + dec_digits(s.end[i]) + 2 +
               (uint64_t)(s.end[i] - s.start[i]) + 1;
}
void launch_fasta_sizes(cudaStream_t st, const DevSampled &s, uint64_t n, const uint64_t *name_off, uint32_t first_index,
                        uint64_t *sizes) {
    if (n) k_fasta_sizes<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(s, n, name_off, first_index, sizes);
}

__global__ void k_fasta_write(DevTranscripts tr, DevSampled s, uint64_t n, const uint8_t *__restrict__ names,
                              const uint64_t *__restrict__ name_off, const uint64_t *__restrict__ rec_off,
                              uint32_t first_index, uint8_t *__restrict__ out) {
    uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t lane = threadIdx.x & 31;
    if (i >= n) return;
    uint32_t t = s.tr[i] - first_index, start = s.start[i], end = s.end[i];
    bool minus = s.minus[i] != 0;
    uint64_t nl = name_off[t + 1] - name_off[t];
    uint8_t *p = out + rec_off[i];
    uint64_t id = s.id[i];
    uint64_t hdr = 4 + dec_digits(id) + 1 + nl + 9 + 1 + 8 + dec_digits(start) + 4 + dec_digits(end) + 2;
    if (lane == 0) {
        uint8_t *q = put_str(p, ">Fg_");
        q = put_dec(q, id);
        *q++ = '_';
        q += nl;  // name copied below by the whole warp
        q = put_str(q, " (Strand ");
        *q++ = minus ? '-' : '+';
        q = put_str(q, " Offset ");
        q = put_dec(q, start);
        q = put_str(q, " -- ");
        q = put_dec(q, end);
        *q++ = ')';
        *q++ = '\n';
    }
    uint8_t *np = p + 4 + dec_digits(id) + 1;
    for (uint64_t b = lane; b < nl; b += 32) np[b] = names[name_off[t] + b];
    const uint8_t *seq = tr.bases + tr.off[t];
    uint8_t *sp = p + hdr;
    uint32_t len = end - start;
    for (uint32_t b = lane; b < len; b += 32) {
        uint8_t c;
        if (!minus) c = seq[start + b];
        else {
            uint8_t x = seq[end - 1 - b];
            switch (x) { case 'A': c = 'T'; break; case 'T': c = 'A'; break; case 'G': c = 'C'; break; case 'C': c = 'G'; break; default: c = 'N'; }
        }
        sp[b] = c;
    }
    if (lane == 0) sp[len] = '\n';
}
void launch_fasta_write(cudaStream_t st, const DevTranscripts &tr, const DevSampled &s, uint64_t n, const uint8_t *names,
                        const uint64_t *name_off, const uint64_t *rec_off, uint32_t first_index, uint8_t *out) {
    if (n) k_fasta_write<<<(unsigned)((n * 32 + 255) / 256), 256, 0, st>>>(tr, s, n, names, name_off, rec_off, first_index, out);
}

}  // namespace rl
