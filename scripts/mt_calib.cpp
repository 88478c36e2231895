#include <immintrin.h>
#include <chrono>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
__attribute__((target("avx512f,avx512bw,avx512vl,avx512dq,bmi2,popcnt")))
int main(){
    const size_t N = 10u<<20; std::vector<uint16_t> w(N+256); for (size_t i=0;i<N+256;i++) w[i]=(uint16_t)(i*2654435761u>>7);
    const __m512i vm=_mm512_set1_epi16(32767), lane=_mm512_set1_epi16(3);
    for (int mode=0; mode<3; mode++){
      auto t0=std::chrono::steady_clock::now(); uint64_t acc=0; uint32_t i=24999;
      for (int rep=0;rep<10;rep++) for (size_t o=0;o+128<=N;o+=128){
        const __m512i vi=_mm512_set1_epi16((short)(mode==2? i : 20000));
        __mmask32 A[4],S[4];
        for (int q=0;q<4;q++){ __m512i v=_mm512_and_si512(_mm512_loadu_si512(&w[o+32*q]),vm); A[q]=_mm512_cmple_epu16_mask(v,vi); S[q]=_mm512_cmple_epu16_mask(_mm512_adds_epu16(v,lane),vi);} 
        uint64_t s0=_cvtmask64_u64(_mm512_kunpackd(S[1],S[0])), s1=_cvtmask64_u64(_mm512_kunpackd(S[3],S[2]));
        uint64_t a0=_cvtmask64_u64(_mm512_kunpackd(A[1],A[0]))&~s0, a1=_cvtmask64_u64(_mm512_kunpackd(A[3],A[2]))&~s1;
        uint32_t tot=__builtin_popcountll(s0)+__builtin_popcountll(s1);
        if (mode>=1) { acc+=tot + (a0|a1); }
        if (mode==2) { i -= tot; if (i<16500) i=24999; }
        else acc ^= s0^s1^a0^a1;
      }
      double dt=std::chrono::duration<double>(std::chrono::steady_clock::now()-t0).count();
      printf("mode %d: %.3f ns/raw (%llu)\n", mode, dt/(10.0*N)*1e9,(unsigned long long)acc+i);
    }
}
