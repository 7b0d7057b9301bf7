/* prints the layout of include/nano_infer_abi.h in the order of oracle/ref_harness.c:orh_abi_layout */
#include <stdio.h>
#include "nano_infer_abi.h"
#define OFF(T, f) printf("%zu ", offsetof(T, f))
int main(void) {
    printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu ", sizeof(LLM_Config), sizeof(LLM_Param), sizeof(FwdBuffer), sizeof(LLM),
           sizeof(Sampler), sizeof(Nano_Context), sizeof(Nano_Session), sizeof(Tokenizer), sizeof(Typed_Tensor), sizeof(LoRA),
           sizeof(Nano_Observation));
    OFF(LLM, config); OFF(LLM, params); OFF(LLM, state); OFF(LLM, arch); OFF(LLM, quant_type); OFF(LLM, group_size);
    OFF(LLM, fd); OFF(LLM, buffer); OFF(LLM, file_size);
    OFF(FwdBuffer, x); OFF(FwdBuffer, xq); OFF(FwdBuffer, q); OFF(FwdBuffer, logits); OFF(FwdBuffer, q0);
    OFF(LLM_Param, token_embedding); OFF(LLM_Param, wq); OFF(LLM_Param, q_norm); OFF(LLM_Param, freq_cis_real); OFF(LLM_Param, token_classifier);
    OFF(Sampler, probindex); OFF(Sampler, repetition_penalty); OFF(Sampler, temperature); OFF(Sampler, top_p); OFF(Sampler, top_k); OFF(Sampler, rng_state);
    OFF(Nano_Context, llm); OFF(Nano_Context, lora); OFF(Nano_Context, tokenizer); OFF(Nano_Context, sampler); OFF(Nano_Context, max_seq_len);
    OFF(Nano_Context, random_seed); OFF(Nano_Context, observation); OFF(Nano_Context, observation_env);
    OFF(Nano_Session, prompt); OFF(Nano_Session, num_prompt_tokens); OFF(Nano_Session, max_seq_len); OFF(Nano_Session, output_ids);
    OFF(Nano_Session, output_count); OFF(Nano_Session, output_text); OFF(Nano_Session, next_token); OFF(Nano_Session, pos);
    OFF(Nano_Session, is_prefilling); OFF(Nano_Session, t_0); OFF(Nano_Session, t_1); OFF(Nano_Session, tps);
    OFF(Tokenizer, vocab_size); OFF(Tokenizer, unicode_charset); OFF(Tokenizer, token_list); OFF(Tokenizer, vocab_trie);
    OFF(Tokenizer, unicode_to_id_map); OFF(Tokenizer, token_to_id_map); OFF(Tokenizer, vocab); OFF(Tokenizer, vocab_scores);
    OFF(Tokenizer, sorted_vocab); OFF(Tokenizer, max_token_length); OFF(Tokenizer, byte_pieces);
    printf("\n");
    return 0;
}
