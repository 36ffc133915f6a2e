// C ABI (include/b200_whisper.h) + the continuous-batching scheduler thread.
#include <math.h>
#include <cmath>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "engine.cuh"

namespace bw {
void engine_build_weight_table(bw_engine* e);
void engine_load_tensor(bw_engine* e, const bw_tensor_desc& t);
void engine_encoder_forward(bw_engine* e, int nb);
void engine_cross_kv(bw_engine* e, int bi, int q);
void engine_window_to_A1(bw_engine* e, const float* logmel, int ld, int n_real, const int* gmax, int seek, int seg, int bi);
void engine_decoder_layers(bw_engine* e, DecGroup& G, int R, int n_groups, int max_group_rows, int n_lrows, const int* row_seq,
                           const int* row_pos, const int* row_tok, const int* row_bpos, const int* grp_first, const int* grp_n,
                           const int* grp_x, const int* lrow_src);
void engine_init_requests(bw_engine* e, const int* init_dev, int n);
void engine_fold_layernorms(bw_engine* e);
void engine_gather_final(bw_engine* e, const int* list_dev, int n, int blob_bytes, unsigned char* out_dev);
}  // namespace bw

using namespace bw;

static thread_local std::string tl_error;

#define BW_API_BEGIN try {
#define BW_API_END                                                     \
  }                                                                    \
  catch (const bw::CudaError& ex) { tl_error = ex.what(); return BW_ERR_CUDA; }      \
  catch (const std::invalid_argument& ex) { tl_error = ex.what(); return BW_ERR_INVALID; } \
  catch (const std::bad_alloc& ex) { tl_error = ex.what(); return BW_ERR_NOMEM; }    \
  catch (const std::exception& ex) { tl_error = ex.what(); return BW_ERR_STATE; }    \
  return BW_OK;

namespace {

struct DeviceGuard {
  int prev = 0;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) BW_CUDA(cudaSetDevice(dev));
  }
  ~DeviceGuard() { cudaSetDevice(prev); }
};

double secs(Clock::time_point a, Clock::time_point b) { return std::chrono::duration<double>(b - a).count(); }

void finish_request(Request* r, int status, const std::string& err) {
  std::lock_guard<std::mutex> g(r->mu);
  r->status = status;
  r->error = err;
  r->done = true;
  r->cv.notify_all();
}

// ---- per-group control block (ints), mirrored host (pinned) / device ----
struct Ctl {
  int *row_seq, *row_pos, *row_tok, *row_bpos, *grp_first, *grp_n, *grp_x, *lrow_src, *srow_lrow, *srow_req, *srow_seq, *act_req,
      *act_first, *ns_lrow, *ns_req;
  int* base = nullptr;
  size_t total = 0;
  // fill counters of the step being built
  int R = 0, NG = 0, LR = 0, SR = 0, NA = 0, NNS = 0, max_grp = 1;
  void layout(int* b, int Rm, int LRm, int Q) {
    base = b;
    int* p = b;
    auto take = [&](size_t n) { int* r = p; p += n; return r; };
    row_seq = take(Rm); row_pos = take(Rm); row_tok = take(Rm); row_bpos = take(Rm);
    grp_first = take(Rm); grp_n = take(Rm); grp_x = take(Rm);
    lrow_src = take(LRm); srow_lrow = take(LRm); srow_req = take(LRm); srow_seq = take(LRm);
    act_req = take(Q); act_first = take(Q); ns_lrow = take(Q); ns_req = take(Q);
    total = (size_t)(p - b);
  }
  void reset() { R = NG = LR = SR = NA = NNS = 0; max_grp = 1; }
};

// H2D of the control block + the whole decoder step (all layers, logits, filters/top-k, beam update) of one group,
// enqueued on the group's stream.  No host synchronisation here.
void enqueue_group_step_eager(bw_engine* e, DecGroup& G, Ctl& c) {
  int* dbase = G.d_ctrl.as<int>();
  auto dev = [&](int* h) { return dbase + (h - c.base); };
  BW_CUDA(cudaMemcpyAsync(dbase, c.base, c.total * 4, cudaMemcpyHostToDevice, G.stream));
  engine_decoder_layers(e, G, c.R, c.NG, c.max_grp, c.LR, dev(c.row_seq), dev(c.row_pos), dev(c.row_tok), dev(c.row_bpos),
                        dev(c.grp_first), dev(c.grp_n), dev(c.grp_x), dev(c.lrow_src));
  const float* logits = G.d_logits.as<float>();
  const int V = e->dims.n_vocab;
  static const bool use_pdl = getenv("B200W_NO_PDL") == nullptr;
  PdlScope pdl(use_pdl && !e->fp32);
  no_speech_prob(logits, V, V, dev(c.ns_lrow), dev(c.ns_req), c.NNS, e->tt.no_speech, e->rs.no_speech_prob, G.stream);
  sample_topk(logits, V, V, dev(c.srow_lrow), dev(c.srow_req), dev(c.srow_seq), c.SR, e->tt, e->rs, e->ss, e->anc_cur,
              G.d_cand_tok.as<int>(), G.d_cand_lp.as<float>(), G.stream);
  beam_update(dev(c.act_req), dev(c.act_first), c.NA, e->tt, e->rs, e->ss, e->anc_cur, e->dims.n_text_ctx, G.d_cand_tok.as<int>(),
              G.d_cand_lp.as<float>(), G.stream);
}

void enqueue_group_step(bw_engine* e, DecGroup& G, Ctl& c) {
  if (c.R == 0) return;
  e->stat_h2d += (long long)c.total * 4;
  static const bool use_graphs = getenv("B200W_NO_GRAPH") == nullptr;
  if (!use_graphs) return enqueue_group_step_eager(e, G, c);
  const StepGraphKey key{c.R, c.NG, c.LR, c.SR, c.NA, c.NNS, c.max_grp, e->anc_cur};
  StepGraph& sg = G.graphs[key];
  if (sg.exec) {
    BW_CUDA(cudaGraphLaunch(sg.exec, G.stream));
    return;
  }
  if (++sg.seen < 3) return enqueue_group_step_eager(e, G, c);  // early sightings also warm every lazy initialisation
  // third sighting of this shape: it is stable enough to pay for a capture (the launch sequence depends only on the key)
  cudaGraph_t graph = nullptr;
  BW_CUDA(cudaStreamBeginCapture(G.stream, cudaStreamCaptureModeThreadLocal));
  try {
    enqueue_group_step_eager(e, G, c);
  } catch (...) {
    cudaStreamEndCapture(G.stream, &graph);
    if (graph) cudaGraphDestroy(graph);
    throw;
  }
  BW_CUDA(cudaStreamEndCapture(G.stream, &graph));
  cudaGraphExec_t exec = nullptr;
  const cudaError_t st = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (st != cudaSuccess) throw CudaError(std::string("cudaGraphInstantiate -> ") + cudaGetErrorString(st));
  if (G.graphs.size() > 96) {  // bound the cache: shapes churn while requests come and go
    for (auto& kv : G.graphs) if (kv.second.exec && !(kv.first == key)) { cudaGraphExecDestroy(kv.second.exec); kv.second.exec = nullptr; kv.second.seen = 0; }
  }
  sg.exec = exec;
  BW_CUDA(cudaGraphLaunch(exec, G.stream));
}

int choose_groups(int n_segments) {
  static const int forced = getenv("B200W_GROUPS") ? atoi(getenv("B200W_GROUPS")) : 0;
  if (forced > 0) return std::min(forced, (int)kMaxGroups);
  (void)n_segments;
  return 1;  // measured: extra groups re-stream the weights and lengthen the step (profiles/r1_notes.md)
}

void release_slots(bw_engine* e, Request* r) {
  if (r->q >= 0) e->free_q.push_back(r->q);
  if (r->first_seq >= 0)
    for (int j = 0; j < r->G; ++j) e->seq_used[r->first_seq + j] = 0;
  r->q = -1;
  r->first_seq = -1;
}

int find_seq_block(bw_engine* e, int G) {
  int run = 0;
  for (int s = 0; s < e->S; ++s) {
    run = e->seq_used[s] ? 0 : run + 1;
    if (run == G) return s - G + 1;
  }
  return -1;
}

// host-side finalisation: BeamSearchDecoder.finalize / GreedyDecoder.finalize + MaximumLikelihoodRanker
void finalize_decode(bw_engine* e, Request* r, const unsigned char* blob) {
  const int n_ctx = e->dims.n_text_ctx;
  const float* fin_score = reinterpret_cast<const float*>(blob);
  const int* fin_pos = reinterpret_cast<const int*>(blob + kMaxFinished * 4);
  const int* fin_slot = fin_pos + kMaxFinished;
  const int* misc = fin_slot + kMaxFinished;  // n_finished
  const float* fmisc = reinterpret_cast<const float*>(misc + 1);  // no_speech_prob, sum_logprob[kMaxBeam]
  const int* tok = reinterpret_cast<const int*>(fmisc + 1 + kMaxBeam);
  const unsigned char* parent = reinterpret_cast<const unsigned char*>(tok + (size_t)n_ctx * kMaxBeam);
  const int sb = (int)r->initial.size();
  const int eot = e->tt.eot;
  auto backtrack = [&](int pos, int slot) {
    std::vector<int> seq;
    for (int t = pos; t >= sb; --t) {
      seq.push_back(tok[(size_t)t * kMaxBeam + slot]);
      slot = parent[(size_t)t * kMaxBeam + slot];
    }
    std::reverse(seq.begin(), seq.end());
    return seq;
  };
  std::vector<std::vector<int>> cand;
  std::vector<float> cand_lp;
  const int last_pos = std::min(r->cur_len - 1, n_ctx - 1);
  if (r->greedy) {
    // GreedyDecoder.finalize: every hypothesis of the group (1, or best_of samples) padded with one EOT
    for (int j = 0; j < r->G; ++j) {
      cand.push_back(backtrack(last_pos, j));
      cand.back().push_back(eot);
      cand_lp.push_back(fmisc[1 + j]);
    }
  } else {
    const int n_fin = misc[0];
    for (int i = 0; i < n_fin; ++i) {
      cand.push_back(backtrack(fin_pos[i], fin_slot[i]));
      cand.back().push_back(eot);
      cand_lp.push_back(fin_score[i]);
    }
    if ((int)cand.size() < r->G) {
      std::vector<int> order(r->G);
      for (int j = 0; j < r->G; ++j) order[j] = j;
      std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return fmisc[1 + a] > fmisc[1 + b]; });
      for (int j : order) {
        std::vector<int> seq = backtrack(last_pos, j);
        seq.push_back(eot);
        bool dup = false;
        for (size_t c = 0; c < cand.size(); ++c)
          if (cand[c] == seq) { cand_lp[c] = fmisc[1 + j]; dup = true; break; }
        if (!dup) { cand.push_back(std::move(seq)); cand_lp.push_back(fmisc[1 + j]); }
        if ((int)cand.size() >= r->G) break;
      }
    }
  }
  int best = 0;
  double best_score = -INFINITY;
  for (size_t c = 0; c < cand.size(); ++c) {
    auto& s = cand[c];
    size_t cut = 0;
    while (cut < s.size() && s[cut] != eot) ++cut;
    s.resize(cut);
    const double len = (double)s.size();
    const double penalty = (r->length_penalty < 0) ? len : pow((5.0 + len) / 6.0, (double)r->length_penalty);
    const double score = (double)cand_lp[c] / penalty;
    if (c == 0 || score > best_score) { best = (int)c; best_score = score; }
  }
  bw_result* o = r->out;
  const auto& sel = cand[best];
  o->n_tokens = (int)std::min(sel.size(), (size_t)BW_MAX_TOKENS);
  for (int i = 0; i < o->n_tokens; ++i) o->tokens[i] = sel[i];
  o->sum_logprob = cand_lp[best];
  o->avg_logprob = (float)((double)cand_lp[best] / (double)(sel.size() + 1));
  o->no_speech_prob = fmisc[0];
  o->n_steps = r->steps;
}

size_t fin_blob_bytes(bw_engine* e) {
  return (size_t)kMaxFinished * 12 + 4 + 4 + kMaxBeam * 4 + (size_t)e->dims.n_text_ctx * kMaxBeam * 4 +
         (size_t)e->dims.n_text_ctx * kMaxBeam + 64;
}

void fail_all(bw_engine* e, std::vector<Request*>& fresh, const std::string& msg, int code) {
  for (Request* r : fresh) { release_slots(e, r); finish_request(r, code, msg); }
  fresh.clear();
  for (Request* r : e->live) { release_slots(e, r); finish_request(r, code, msg); }
  e->live.clear();
}

void admit_batch(bw_engine* e, std::vector<Request*>& fresh) {
  const auto& d = e->dims;
  const int nb = (int)fresh.size();
  for (int i = 0; i < nb; ++i) {
    Request* r = fresh[i];
    r->batch_index = i;
    r->t_admit = Clock::now();
    if (r->kind == REQ_LOGITS) {
      const size_t n = (size_t)d.n_mels * 3000;
      float* tmp = e->staging.as<float>();
      BW_CUDA(cudaMemcpyAsync(tmp, r->host_mel, n * 4, cudaMemcpyHostToDevice, e->stream));
      engine_window_to_A1(e, tmp, 3000, 3000, nullptr, 0, 3000, i);
      BW_CUDA(cudaStreamSynchronize(e->stream));  // staging is reused by the next request
    } else {
      bw_call* c = r->call;
      BW_CUDA(cudaStreamWaitEvent(e->stream, c->mel_done, 0));
      const int seg = std::max(0, std::min(3000, c->content_frames - r->seek));
      engine_window_to_A1(e, c->buf.logmel, c->buf.ld, c->n_real, c->buf.gmax, r->seek, seg, i);
    }
  }
  engine_encoder_forward(e, nb);
  for (int i = 0; i < nb; ++i) engine_cross_kv(e, i, fresh[i]->q);
  for (int i = 0; i < nb; ++i) {
    Request* r = fresh[i];
    int* rec = e->h_init + i * kInitRecInts;
    const int n_init = (int)r->initial.size();
    rec[0] = r->q; rec[1] = r->G; rec[2] = r->greedy; rec[3] = n_init; rec[4] = n_init; rec[5] = r->first_seq;
    rec[6] = r->without_ts; rec[7] = r->suppress_blank; rec[8] = r->max_initial_ts;
    rec[9] = std::max(1, (int)lround((double)r->G * (double)r->patience));
    rec[10] = r->initial.back();
    memcpy(&rec[11], &r->temperature, 4);
    rec[12] = (int)(unsigned int)(r->seed & 0xffffffffull); rec[13] = (int)(unsigned int)(r->seed >> 32);
    rec[14] = rec[15] = 0;
    r->cur_len = n_init;
    r->steps = 0;
    r->prefilled = false;
  }
  int* init_dev = e->d_init.as<int>();
  BW_CUDA(cudaMemcpyAsync(init_dev, e->h_init, (size_t)nb * kInitRecInts * 4, cudaMemcpyHostToDevice, e->stream));
  engine_init_requests(e, init_dev, nb);
  // the pinned control block is rewritten by the next step: make sure the copy has been consumed
  BW_CUDA(cudaStreamSynchronize(e->stream));
  e->stat_windows += nb;
  e->stat_enc_batches += 1;
  e->stat_h2d += (long long)nb * kInitRecInts * 4;
  for (Request* r : fresh) { r->t_encoded = Clock::now(); e->live.push_back(r); }
  fresh.clear();
}

void decode_step(bw_engine* e, Ctl* ctls) {
  const auto& d = e->dims;
  struct Special { Request* r; int grp; int lrow0; int n; };
  std::vector<Special> lang_reqs, logit_reqs;
  const int ng = choose_groups((int)e->live.size());
  for (int g = 0; g < ng; ++g) ctls[g].reset();
  for (Request* r : e->live) {
    int gi = 0;  // least-loaded group (rows)
    for (int g = 1; g < ng; ++g) if (ctls[g].R < ctls[gi].R) gi = g;
    Ctl& ctl = ctls[gi];
    int &R = ctl.R, &NG = ctl.NG, &LR = ctl.LR, &SR = ctl.SR, &NA = ctl.NA, &NNS = ctl.NNS, &max_grp = ctl.max_grp;
    if (!r->prefilled) {
      const int n_init = (int)r->initial.size();
      const int row0 = R;
      for (int t = 0; t < n_init; ++t) {
        ctl.row_seq[R] = r->first_seq; ctl.row_pos[R] = t; ctl.row_tok[R] = r->initial[t]; ctl.row_bpos[R] = 0;
        ++R;
      }
      for (int t = 0; t < n_init; t += 8) {
        ctl.grp_first[NG] = row0 + t; ctl.grp_n[NG] = std::min(8, n_init - t); ctl.grp_x[NG] = r->q;
        max_grp = std::max(max_grp, ctl.grp_n[NG]);
        ++NG;
      }
      if (r->kind == REQ_LANG) {
        ctl.lrow_src[LR] = row0;
        lang_reqs.push_back({r, gi, LR, 1});
        ++LR;
      } else if (r->kind == REQ_LOGITS) {
        logit_reqs.push_back({r, gi, LR, n_init});
        for (int t = 0; t < n_init; ++t) ctl.lrow_src[LR++] = row0 + t;
      } else {
        const int last_row = row0 + n_init - 1;
        int sot_lrow;
        if (r->sot_index != n_init - 1) {
          ctl.lrow_src[LR] = row0 + r->sot_index;
          sot_lrow = LR++;
        } else sot_lrow = LR;
        ctl.lrow_src[LR] = last_row;
        ctl.ns_lrow[NNS] = sot_lrow; ctl.ns_req[NNS] = r->q; ++NNS;
        ctl.act_req[NA] = r->q; ctl.act_first[NA] = SR; ++NA;
        ctl.srow_lrow[SR] = LR; ctl.srow_req[SR] = r->q; ctl.srow_seq[SR] = r->first_seq; ++SR;
        ++LR;
      }
    } else {
      ctl.grp_first[NG] = R; ctl.grp_n[NG] = r->G; ctl.grp_x[NG] = r->q;
      max_grp = std::max(max_grp, r->G);
      ++NG;
      ctl.act_req[NA] = r->q; ctl.act_first[NA] = SR; ++NA;
      for (int j = 0; j < r->G; ++j) {
        ctl.row_seq[R] = r->first_seq + j; ctl.row_pos[R] = r->cur_len - 1; ctl.row_tok[R] = -1; ctl.row_bpos[R] = r->cur_len - 1;
        ctl.lrow_src[LR] = R;
        ctl.srow_lrow[SR] = LR; ctl.srow_req[SR] = r->q; ctl.srow_seq[SR] = r->first_seq + j;
        ++R; ++LR; ++SR;
      }
    }
  }
  int total_rows = 0;
  for (int g = 0; g < ng; ++g) { enqueue_group_step(e, e->grp[g], ctls[g]); total_rows += ctls[g].R; }
  const int V = d.n_vocab;
  for (auto& s : lang_reqs) {
    DecGroup& G = e->grp[s.grp];
    language_probs(G.d_logits.as<float>() + (size_t)s.lrow0 * V, V, e->tt.first_language_token, e->tt.num_languages,
                   e->d_lang_probs.as<float>(), e->d_lang_arg.as<int>(), G.stream);
    BW_CUDA(cudaMemcpyAsync(s.r->lang_out->probs, e->d_lang_probs.p, (size_t)e->tt.num_languages * 4, cudaMemcpyDeviceToHost, G.stream));
    BW_CUDA(cudaMemcpyAsync(&s.r->lang_out->language_token, e->d_lang_arg.p, 4, cudaMemcpyDeviceToHost, G.stream));
    BW_CUDA(cudaStreamSynchronize(G.stream));
    s.r->lang_out->n_languages = e->tt.num_languages;
  }
  for (auto& s : logit_reqs)
    BW_CUDA(cudaMemcpyAsync(s.r->logits_out, e->grp[s.grp].d_logits.as<float>() + (size_t)s.lrow0 * V, (size_t)s.n * V * 4,
                            cudaMemcpyDeviceToHost, e->grp[s.grp].stream));
  for (int g = 0; g < ng; ++g) BW_CUDA(cudaStreamSynchronize(e->grp[g].stream));
  BW_CUDA(cudaMemcpyAsync(e->h_flags, e->rs.completed, (size_t)e->Q * 4, cudaMemcpyDeviceToHost, e->stream));
  BW_CUDA(cudaStreamSynchronize(e->stream));
  e->stat_d2h += (long long)e->Q * 4;
  e->anc_cur ^= 1;
  e->stat_steps += 1;
  e->stat_rows += total_rows;

  // bookkeeping + completion
  std::vector<Request*> still, done;
  for (Request* r : e->live) {
    if (r->kind != REQ_DECODE) { done.push_back(r); continue; }
    r->prefilled = true;
    r->cur_len += 1;
    r->steps += 1;
    if (e->h_flags[r->q] || r->steps >= r->sample_len || r->cur_len > d.n_text_ctx) done.push_back(r);
    else still.push_back(r);
  }
  if (!done.empty()) {
    const size_t blob = fin_blob_bytes(e);
    size_t nd = 0;
    for (Request* r : done) {  // (q, first sequence, hypotheses) of every finished decode; h_init is idle during a step
      if (r->kind != REQ_DECODE) continue;
      int* rec = e->h_init + nd * 3;
      rec[0] = r->q; rec[1] = r->first_seq; rec[2] = r->G;
      ++nd;
    }
    if (nd > 0) {
      BW_CUDA(cudaMemcpyAsync(e->d_init.p, e->h_init, nd * 12, cudaMemcpyHostToDevice, e->stream));
      engine_gather_final(e, e->d_init.as<int>(), (int)nd, (int)blob, e->d_fin.as<unsigned char>());
      BW_CUDA(cudaMemcpyAsync(e->h_fin, e->d_fin.p, nd * blob, cudaMemcpyDeviceToHost, e->stream));
      e->stat_d2h += (long long)(nd * blob);
    }
    BW_CUDA(cudaStreamSynchronize(e->stream));
    nd = 0;
    const auto now = Clock::now();
    for (Request* r : done) {
      if (r->kind == REQ_DECODE) {
        finalize_decode(e, r, e->h_fin + nd * blob);
        r->out->t_queue = (float)secs(r->t_submit, r->t_admit);
        r->out->t_encode = (float)secs(r->t_admit, r->t_encoded);
        r->out->t_decode = (float)secs(r->t_encoded, now);
        ++nd;
      }
      {
        std::lock_guard<std::mutex> g(e->q_mu);
        release_slots(e, r);
      }
      finish_request(r, BW_OK, "");
    }
  }
  e->live.swap(still);
}

void scheduler_main(bw_engine* e) {
  cudaSetDevice(e->device);
  Ctl ctls[kMaxGroups];
  for (int g = 0; g < kMaxGroups; ++g) ctls[g].layout(e->grp[g].h_ctrl, e->R_max, e->LR_max, e->Q);
  const char* wenv = getenv("B200W_BATCH_WINDOW_US");
  const int window_us = wenv ? atoi(wenv) : 300;
  for (;;) {
    std::vector<Request*> fresh;
    {
      std::unique_lock<std::mutex> lk(e->q_mu);
      e->q_cv.wait(lk, [&] { return e->stop || !e->pending.empty() || !e->live.empty(); });
      if (e->stop) {
        std::vector<Request*> all(e->pending.begin(), e->pending.end());
        e->pending.clear();
        lk.unlock();
        fail_all(e, all, "engine destroyed", BW_ERR_STATE);
        return;
      }
      if (e->live.empty() && !e->pending.empty() && (int)e->pending.size() < e->Be && window_us > 0) {
        // short batching window so that simultaneous callers share one encoder launch
        e->q_cv.wait_for(lk, std::chrono::microseconds(window_us), [&] { return e->stop || (int)e->pending.size() >= e->Be; });
      }
      int rows = 0, lrows = 0;
      for (Request* r : e->live) { rows += r->G; lrows += r->G; }
      while (!e->pending.empty() && (int)fresh.size() < e->Be) {
        Request* r = e->pending.front();
        const int n_init = (int)r->initial.size();
        const int need_l = (r->kind == REQ_LOGITS) ? n_init : 2;
        if (e->free_q.empty() || rows + n_init > e->R_max || lrows + need_l > e->LR_max) break;
        const int fs = find_seq_block(e, r->G);
        if (fs < 0) break;
        r->q = e->free_q.back();
        e->free_q.pop_back();
        r->first_seq = fs;
        for (int j = 0; j < r->G; ++j) e->seq_used[fs + j] = 1;
        rows += n_init;
        lrows += need_l;
        fresh.push_back(r);
        e->pending.pop_front();
      }
    }
    try {
      std::lock_guard<std::mutex> g(e->gpu_mu);
      if (!fresh.empty()) admit_batch(e, fresh);
      if (!e->live.empty()) decode_step(e, ctls);
    } catch (const std::exception& ex) {
      std::lock_guard<std::mutex> g(e->q_mu);
      fail_all(e, fresh, ex.what(), BW_ERR_CUDA);
    }
  }
}

int submit_and_wait(bw_engine* e, Request& r) {
  if (e->state != 1 || !e->sched.joinable()) { tl_error = "engine not finalized or scheduler disabled"; return BW_ERR_STATE; }
  r.t_submit = Clock::now();
  {
    std::lock_guard<std::mutex> g(e->q_mu);
    if (e->stop) { tl_error = "engine stopping"; return BW_ERR_STATE; }
    e->pending.push_back(&r);
  }
  e->q_cv.notify_all();
  std::unique_lock<std::mutex> lk(r.mu);
  r.cv.wait(lk, [&] { return r.done; });
  if (r.status != BW_OK) tl_error = r.error;
  return r.status;
}

void compute_call_mel(bw_engine* e, bw_call* c, const float* pcm, long long n) {
  const unsigned idx = e->front_rr.fetch_add(1) % bw_engine::kFrontStreams;
  std::lock_guard<std::mutex> g(e->front_mu[idx]);
  cudaStream_t st = e->front[idx];
  BW_CUDA(cudaMemcpyAsync(c->buf.pcm, pcm, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  mel_power(c->buf.pcm, n, 480000, e->mel_tables.as<float>(), e->mel_filters.as<float>(), e->mel_ranges.as<int2>(), e->dims.n_mels,
            c->buf.logmel, c->buf.ld, c->n_real, c->total_frames, c->buf.gmax, st);
  BW_CUDA(cudaEventRecord(c->mel_done, st));
  e->stat_h2d += n * 4;
}

// raw ingest: H2D of the int16 samples, conversion (+ resampling) into the call's f32 PCM buffer, then the log-mel
const Resampler* find_resampler(bw_engine* e, int sample_rate) {
  if (sample_rate == 16000) return nullptr;
  std::lock_guard<std::mutex> g(e->resampler_mu);
  auto it = e->resamplers.find(sample_rate);
  if (it == e->resamplers.end())
    throw std::invalid_argument("no resampler registered for " + std::to_string(sample_rate) + " Hz (bw_engine_set_resampler)");
  return it->second.get();
}
long long resampled_length(const Resampler* r, long long n) {
  return r ? ((long long)r->nw * n + r->orig - 1) / r->orig : n;
}
// enqueue on `st`: pcm16 (host) -> staging -> f32 PCM at 16 kHz in `dst`
void ingest_pcm16(bw_engine* e, int front_idx, cudaStream_t st, const int16_t* pcm, long long n, const Resampler* r, float* dst, long long n16) {
  DevBuf& stage = e->front_pcm16[front_idx];
  if (stage.bytes < (size_t)n * 2 + 16) {
    BW_CUDA(cudaStreamSynchronize(st));  // earlier users of the old staging buffer
    stage.alloc(std::max((size_t)n * 2 + 16, (size_t)4 << 20));
  }
  BW_CUDA(cudaMemcpyAsync(stage.p, pcm, (size_t)n * 2, cudaMemcpyHostToDevice, st));
  if (r) pcm16_resample(stage.as<int16_t>(), n, r->taps.as<float>(), r->ranges.as<int2>(), r->orig, r->nw, r->K, r->width, dst, n16, st);
  else pcm16_to_f32(stage.as<int16_t>(), n, dst, st);
  e->stat_h2d += n * 2;
}

}  // namespace

// ================================================================================================
extern "C" {

const char* bw_last_error(void) { return tl_error.c_str(); }
int bw_version(void) { return 100; }
int bw_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int bw_engine_create(const bw_model_dims* dims, const bw_engine_config* cfg, bw_engine** out) {
  BW_API_BEGIN
  BW_CHECK(dims && cfg && out, "null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { tl_error = "no CUDA device visible"; return BW_ERR_NO_DEVICE; }
  BW_CHECK(cfg->cuda_device >= 0 && cfg->cuda_device < ndev, "cuda_device out of range");
  cudaDeviceProp prop;
  BW_CUDA(cudaGetDeviceProperties(&prop, cfg->cuda_device));
  if (prop.major != 10) { tl_error = std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) + ", this library is built for sm_100a only"; return BW_ERR_NO_DEVICE; }
  BW_CHECK(dims->n_audio_state == dims->n_audio_head * 64 && dims->n_text_state == dims->n_text_head * 64, "head dim must be 64");
  BW_CHECK(dims->n_audio_state == dims->n_text_state, "encoder/decoder width must match");
  BW_CHECK(dims->n_audio_ctx == 1500 && dims->n_text_ctx <= BW_MAX_TOKENS && dims->n_text_ctx >= 8, "unsupported context sizes");
  BW_CHECK(dims->n_audio_state <= 1280 && dims->n_audio_state % 64 == 0, "n_state must be <= 1280");
  BW_CHECK(dims->n_mels > 0 && dims->n_mels <= 128 && (dims->n_mels * 3) % 8 == 0, "unsupported n_mels");
  BW_CHECK(dims->n_vocab > 50000 && dims->n_vocab < 60000, "unsupported vocabulary size");
  DeviceGuard dg(cfg->cuda_device);
  auto e = std::make_unique<bw_engine>();
  e->dims = *dims;
  e->cfg = *cfg;
  e->device = cfg->cuda_device;
  e->fp32 = cfg->compute == BW_COMPUTE_FP32;
  e->force_simt = (cfg->flags & BW_FLAG_FORCE_SIMT_GEMM) != 0;
  e->fuse_ln = !e->fp32 && !e->force_simt && (dims->n_text_state % 64) == 0 && getenv("B200W_NO_LN_FUSION") == nullptr;
  BW_CUDA(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  BW_CUDA(cudaEventCreateWithFlags(&e->enc_fork, cudaEventDisableTiming));
  for (int i = 0; i < bw_engine::kEncStreams; ++i) {
    BW_CUDA(cudaStreamCreateWithFlags(&e->enc_streams[i], cudaStreamNonBlocking));
    BW_CUDA(cudaEventCreateWithFlags(&e->enc_join[i], cudaEventDisableTiming));
  }
  for (auto& s : e->front) BW_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  e->staging.alloc((size_t)64 << 20);
  engine_build_weight_table(e.get());
  e->loaded.assign(e->expected_names.size() + 1, 0);
  // mel constant tables
  std::vector<float> tab(mel_tables_floats());
  mel_fill_tables(tab.data());
  e->mel_tables.alloc(tab.size() * 4);
  BW_CUDA(cudaMemcpy(e->mel_tables.p, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice));
  *out = e.release();
  BW_API_END
}

int bw_engine_load_weights(bw_engine* e, const bw_tensor_desc* tensors, int32_t n) {
  BW_API_BEGIN
  BW_CHECK(e && tensors, "null argument");
  BW_CHECK(e->state == 0, "weights must be loaded before bw_engine_finalize");
  DeviceGuard dg(e->device);
  for (int i = 0; i < n; ++i) engine_load_tensor(e, tensors[i]);
  BW_API_END
}

int bw_engine_set_tables(bw_engine* e, const bw_token_tables* t) {
  BW_API_BEGIN
  BW_CHECK(e && t, "null argument");
  DeviceGuard dg(e->device);
  const int V = e->dims.n_vocab;
  TokenTables& tt = e->tt;
  tt.eot = t->eot; tt.sot = t->sot; tt.sot_prev = t->sot_prev; tt.sot_lm = t->sot_lm; tt.no_speech = t->no_speech;
  tt.no_timestamps = t->no_timestamps; tt.timestamp_begin = t->timestamp_begin; tt.translate = t->translate;
  tt.transcribe = t->transcribe; tt.first_language_token = t->first_language_token; tt.num_languages = t->num_languages;
  BW_CHECK(t->n_blank >= 0 && t->n_blank <= 4, "at most 4 blank ids");
  BW_CHECK(t->timestamp_begin > 0 && t->timestamp_begin < V && t->eot < t->timestamp_begin, "bad token layout");
  BW_CHECK(t->num_languages <= 128, "too many languages");
  tt.n_blank = t->n_blank;
  for (int i = 0; i < 4; ++i) tt.blank[i] = (i < t->n_blank) ? t->blank[i] : -1;
  std::vector<unsigned int> bits((V + 31) / 32, 0u);
  for (int i = 0; i < t->n_suppress; ++i) {
    BW_CHECK(t->suppress[i] >= 0 && t->suppress[i] < V, "suppress id out of range");
    bits[t->suppress[i] >> 5] |= 1u << (t->suppress[i] & 31);
  }
  e->suppress_bits.alloc(bits.size() * 4);
  BW_CUDA(cudaMemcpy(e->suppress_bits.p, bits.data(), bits.size() * 4, cudaMemcpyHostToDevice));
  tt.suppress_bits = e->suppress_bits.as<unsigned int>();
  e->tables_set = true;
  BW_API_END
}

int bw_engine_set_mel_filters(bw_engine* e, const float* filters) {
  BW_API_BEGIN
  BW_CHECK(e && filters, "null argument");
  DeviceGuard dg(e->device);
  const int nm = e->dims.n_mels;
  std::vector<int2> ranges(nm);
  for (int m = 0; m < nm; ++m) {
    int lo = 201, hi = 0;
    for (int k = 0; k < 201; ++k)
      if (filters[m * 201 + k] != 0.f) { lo = std::min(lo, k); hi = std::max(hi, k + 1); }
    if (lo > hi) { lo = 0; hi = 0; }
    ranges[m] = make_int2(lo, hi);
  }
  e->mel_filters.alloc((size_t)nm * 201 * 4);
  e->mel_ranges.alloc((size_t)nm * sizeof(int2));
  BW_CUDA(cudaMemcpy(e->mel_filters.p, filters, (size_t)nm * 201 * 4, cudaMemcpyHostToDevice));
  BW_CUDA(cudaMemcpy(e->mel_ranges.p, ranges.data(), (size_t)nm * sizeof(int2), cudaMemcpyHostToDevice));
  e->filters_set = true;
  BW_API_END
}

int bw_engine_finalize(bw_engine* e) {
  BW_API_BEGIN
  BW_CHECK(e, "null argument");
  BW_CHECK(e->state == 0, "already finalized");
  BW_CHECK(e->tables_set && e->filters_set, "bw_engine_set_tables and bw_engine_set_mel_filters must be called first");
  DeviceGuard dg(e->device);
  const auto& d = e->dims;
  for (size_t i = 0; i < e->expected_names.size(); ++i)
    if (!e->loaded[i]) throw std::invalid_argument("missing weight tensor: " + e->expected_names[i]);
  if (!e->loaded.back()) {  // encoder.positional_embedding = sinusoids(n_audio_ctx, d)
    const int ch = d.n_audio_state, len = d.n_audio_ctx;
    std::vector<float> pe((size_t)len * ch);
    const float inc = (float)(log(10000.0) / (ch / 2 - 1));
    for (int t = 0; t < len; ++t)
      for (int i = 0; i < ch / 2; ++i) {
        const float inv = expf(-inc * (float)i);
        const float a = (float)t * inv;
        pe[(size_t)t * ch + i] = sinf(a);
        pe[(size_t)t * ch + ch / 2 + i] = cosf(a);
      }
    BW_CUDA(cudaMemcpy(e->w.enc_pos, pe.data(), pe.size() * 4, cudaMemcpyHostToDevice));
  }
  engine_fold_layernorms(e);
  const size_t ts = e->fp32 ? 4 : 2;
  const size_t dm = d.n_audio_state;
  const size_t cross_slot = (size_t)d.n_text_layer * d.n_audio_ctx * 2 * dm * ts;
  const size_t unit = (size_t)d.n_text_layer * 2 * d.n_text_ctx * dm * ts;
  int Q = e->cfg.max_segments > 0 ? e->cfg.max_segments : 64;
  int S = e->cfg.max_sequences > 0 ? e->cfg.max_sequences : std::max(2 * Q, kMaxBeam);
  int Be = e->cfg.max_encoder_batch > 0 ? e->cfg.max_encoder_batch : 8;
  size_t free_b = 0, total_b = 0;
  BW_CUDA(cudaMemGetInfo(&free_b, &total_b));
  const size_t enc_per = (size_t)(3000 * 3 * d.n_mels + 3000 * dm + 1500 * 3 * dm + 1500 * dm * 2 + 1500 * 3 * dm + 1500 * dm +
                                  1500 * 4 * dm + 1500 * dm) * 4;  // upper bound (fp32 sizes)
  for (int guard = 0; guard < 64; ++guard) {
    if (!((double)Q * cross_slot + (double)S * unit + (double)Be * enc_per > 0.80 * (double)free_b && (Q > 1 || S > kMaxBeam || Be > 1))) break;
    if (Q > 1) Q = std::max(1, Q * 3 / 4);
    S = std::max(kMaxBeam, std::min(S, std::max(2 * Q, kMaxBeam)));
    if (Be > 1) Be = std::max(1, Be / 2);
  }
  while (false && (double)Q * cross_slot + (double)S * unit + (double)Be * enc_per > 0.80 * (double)free_b && (Q > 1 || S > kMaxBeam || Be > 1)) {
    if (Q > 1) Q = std::max(1, Q * 3 / 4);
    S = std::max(kMaxBeam, std::min(S, std::max(2 * Q, kMaxBeam)));
    if (Be > 1 && (double)Be * enc_per > 0.2 * (double)free_b) Be = std::max(1, Be / 2);
    if (Q == 1 && Be == 1) break;
  }
  Be = std::min(Be, Q);
  e->Q = Q; e->S = S; e->Be = Be;
  e->R_max = S + 512;
  e->LR_max = S + Q + BW_MAX_TOKENS;
  e->cross_cache.alloc((size_t)Q * cross_slot);
  e->self_pool.alloc((size_t)S * unit);
  // encoder activations
  e->A1.alloc((size_t)Be * 3000 * 3 * d.n_mels * ts);
  e->y1.alloc((size_t)Be * 3000 * dm * ts);
  e->A2.alloc((size_t)Be * 1500 * 3 * dm * ts);
  e->enc_x.alloc((size_t)Be * 1500 * dm * 4);
  e->enc_xn.alloc((size_t)Be * 1500 * dm * ts);
  e->enc_qkv.alloc((size_t)Be * 1500 * 3 * dm * ts);
  e->enc_att.alloc((size_t)Be * 1500 * dm * ts);
  e->enc_h.alloc((size_t)Be * 1500 * 4 * dm * ts);
  e->enc_out.alloc((size_t)Be * 1500 * dm * ts);
  // decoder activations
  const size_t R = e->R_max, LR = e->LR_max;
  for (int gi = 0; gi < kMaxGroups; ++gi) {
    DecGroup& G = e->grp[gi];
    BW_CUDA(cudaStreamCreateWithFlags(&G.stream, cudaStreamNonBlocking));
    G.d_x.alloc(R * dm * 4);
    G.d_xn.alloc(R * dm * ts); G.d_qkv.alloc(R * 3 * dm * 4); G.d_att.alloc(R * dm * ts); G.d_q.alloc(R * dm * 4);
    G.d_h.alloc(R * 4 * dm * ts); G.d_lnrows.alloc(LR * dm * ts);
    if (e->fuse_ln) {
      G.d_xb.alloc(R * dm * 2);
      G.d_lnst.alloc(R * (dm / 64) * sizeof(float2));
      BW_CUDA(cudaMemset(G.d_xb.p, 0, G.d_xb.bytes));
      BW_CUDA(cudaMemset(G.d_lnst.p, 0, G.d_lnst.bytes));
    }
    G.d_logits.alloc(LR * (size_t)d.n_vocab * 4);
    G.d_ws.alloc(dec_cross_workspace_floats((int)R, d.n_text_head) * 4);
    G.d_cand_tok.alloc(LR * kMaxCand * 4); G.d_cand_lp.alloc(LR * kMaxCand * 4);
    // rows past the live count are read (and discarded) by the GEMMs' TMA boxes: keep them finite
    for (DevBuf* b : {&G.d_xn, &G.d_qkv, &G.d_att, &G.d_q, &G.d_h, &G.d_lnrows}) BW_CUDA(cudaMemset(b->p, 0, b->bytes));
  }
  e->d_lang_probs.alloc(128 * 4); e->d_lang_arg.alloc(4);
  BW_CUDA(cudaMemset(e->self_pool.p, 0, e->self_pool.bytes));
  // decoder state
  const size_t n_ctx = d.n_text_ctx;
  e->st_int.alloc(((size_t)Q * 13 + (size_t)Q * kMaxFinished * 2 + (size_t)S * 4) * 4);
  e->st_float.alloc(((size_t)Q * kMaxFinished + 2 * (size_t)Q + S) * 4);
  e->st_anc0.alloc((size_t)S * n_ctx); e->st_anc1.alloc((size_t)S * n_ctx);
  e->st_tok.alloc((size_t)Q * n_ctx * kMaxBeam * 4); e->st_parent.alloc((size_t)Q * n_ctx * kMaxBeam);
  for (DevBuf* b : {&e->st_int, &e->st_float, &e->st_anc0, &e->st_anc1, &e->st_tok, &e->st_parent}) BW_CUDA(cudaMemset(b->p, 0, b->bytes));
  {
    int* p = e->st_int.as<int>();
    auto take = [&](size_t n) { int* r = p; p += n; return r; };
    ReqState& rs = e->rs;
    rs.n_beam = take(Q); rs.greedy = take(Q); rs.sample_begin = take(Q); rs.cur_len = take(Q); rs.first_seq = take(Q);
    rs.without_ts = take(Q); rs.suppress_blank = take(Q); rs.max_initial_ts = take(Q); rs.max_candidates = take(Q);
    rs.n_finished = take(Q); rs.completed = take(Q);
    rs.seed_lo = reinterpret_cast<unsigned int*>(take(Q)); rs.seed_hi = reinterpret_cast<unsigned int*>(take(Q));
    rs.fin_pos = take((size_t)Q * kMaxFinished); rs.fin_slot = take((size_t)Q * kMaxFinished);
    SeqState& ss = e->ss;
    ss.next_tok = take(S); ss.prev_tok = take(S); ss.last_ts = take(S); ss.seq_first = take(S);
    float* f = e->st_float.as<float>();
    rs.fin_score = f; f += (size_t)Q * kMaxFinished;
    rs.no_speech_prob = f; f += Q;
    rs.temperature = f; f += Q;
    ss.sum_logprob = f;
    ss.anc[0] = e->st_anc0.as<unsigned char>(); ss.anc[1] = e->st_anc1.as<unsigned char>();
    rs.tok = e->st_tok.as<int>(); rs.parent = e->st_parent.as<unsigned char>();
  }
  // control blocks (one per group) + admission records
  {
    Ctl probe;
    probe.layout(nullptr, e->R_max, e->LR_max, Q);
    e->ctrl_ints = probe.total;
    for (int gi = 0; gi < kMaxGroups; ++gi) {
      BW_CUDA(cudaMallocHost(reinterpret_cast<void**>(&e->grp[gi].h_ctrl), e->ctrl_ints * 4));
      e->grp[gi].d_ctrl.alloc(e->ctrl_ints * 4);
    }
    BW_CUDA(cudaMallocHost(reinterpret_cast<void**>(&e->h_init), (size_t)Q * kInitRecInts * 4));
    e->d_init.alloc((size_t)Q * kInitRecInts * 4);
    BW_CUDA(cudaMallocHost(reinterpret_cast<void**>(&e->h_flags), (size_t)Q * 4));
    e->h_fin_bytes = fin_blob_bytes(e) * (size_t)Q;
    BW_CUDA(cudaMallocHost(reinterpret_cast<void**>(&e->h_fin), e->h_fin_bytes));
    e->d_fin.alloc(e->h_fin_bytes);
  }
  // call buffers: sized for 30 s of audio; longer calls allocate on demand
  e->call_pcm_cap = 480000 + 1600;
  e->call_ld = 3008;
  const int n_calls = 2 * Q + 8;
  for (int i = 0; i < n_calls; ++i) {
    CallBuf b;
    BW_CUDA(cudaMalloc(reinterpret_cast<void**>(&b.pcm), (size_t)e->call_pcm_cap * 4));
    BW_CUDA(cudaMalloc(reinterpret_cast<void**>(&b.logmel), (size_t)d.n_mels * e->call_ld * 4));
    BW_CUDA(cudaMalloc(reinterpret_cast<void**>(&b.gmax), 4));
    b.pcm_cap = e->call_pcm_cap; b.ld = e->call_ld; b.pooled = true;
    e->call_pool.push_back(b);
  }
  e->free_q.clear();
  for (int q = Q - 1; q >= 0; --q) e->free_q.push_back(q);
  e->seq_used.assign(S, 0);
  BW_CUDA(cudaDeviceSynchronize());
  e->state = 1;
  if (!(e->cfg.flags & BW_FLAG_NO_SCHEDULER)) e->sched = std::thread(scheduler_main, e);
  BW_API_END
}

int bw_engine_retain(bw_engine* e) {
  if (!e) return BW_ERR_INVALID;
  e->refs.fetch_add(1);
  return BW_OK;
}

int bw_engine_destroy(bw_engine* e) {
  BW_API_BEGIN
  if (!e) return BW_OK;
  if (e->refs.fetch_sub(1) > 1) return BW_OK;
  {
    std::lock_guard<std::mutex> g(e->q_mu);
    e->stop = true;
  }
  e->q_cv.notify_all();
  if (e->sched.joinable()) e->sched.join();
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  for (auto& b : e->call_pool) { cudaFree(b.pcm); cudaFree(b.logmel); cudaFree(b.gmax); }
  for (auto& G : e->grp) {
    for (auto& kv : G.graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    if (G.h_ctrl) cudaFreeHost(G.h_ctrl);
    if (G.stream) cudaStreamDestroy(G.stream);
  }
  if (e->h_init) cudaFreeHost(e->h_init);
  if (e->h_flags) cudaFreeHost(e->h_flags);
  if (e->h_fin) cudaFreeHost(e->h_fin);
  for (auto& s : e->front) if (s) cudaStreamDestroy(s);
  if (e->enc_fork) cudaEventDestroy(e->enc_fork);
  for (int i = 0; i < bw_engine::kEncStreams; ++i) {
    if (e->enc_join[i]) cudaEventDestroy(e->enc_join[i]);
    if (e->enc_streams[i]) cudaStreamDestroy(e->enc_streams[i]);
  }
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
  BW_API_END
}

int bw_engine_stats(bw_engine* e, int64_t* out, int32_t n) {
  BW_API_BEGIN
  BW_CHECK(e && out, "null argument");
  int64_t v[BW_STAT_COUNT] = {0};
  v[BW_STAT_KERNEL_LAUNCHES] = g_kernel_launches.load();
  v[BW_STAT_DECODE_STEPS] = e->stat_steps; v[BW_STAT_ROWS] = e->stat_rows; v[BW_STAT_WINDOWS] = e->stat_windows;
  v[BW_STAT_MAX_SEGMENTS] = e->Q; v[BW_STAT_MAX_SEQUENCES] = e->S; v[BW_STAT_ENCODER_BATCHES] = e->stat_enc_batches;
  v[BW_STAT_H2D_BYTES] = e->stat_h2d; v[BW_STAT_D2H_BYTES] = e->stat_d2h;
  for (int i = 0; i < n && i < BW_STAT_COUNT; ++i) out[i] = v[i];
  BW_API_END
}

// ---- calls ----
int bw_call_open(bw_engine* e, const float* pcm, int64_t n_samples, bw_call** out) {
  BW_API_BEGIN
  BW_CHECK(e && pcm && out, "null argument");
  BW_CHECK(e->state == 1, "engine not finalized");
  BW_CHECK(n_samples >= 1, "empty audio");
  DeviceGuard dg(e->device);
  auto c = std::make_unique<bw_call>();
  c->eng = e;
  c->n_samples = n_samples;
  c->total_frames = (int)((n_samples + 480000) / 160);
  c->content_frames = c->total_frames - 3000;
  c->n_real = (int)std::min<long long>(c->total_frames, (n_samples + 200 + 159) / 160);
  {
    std::lock_guard<std::mutex> g(e->call_mu);
    if (n_samples <= e->call_pcm_cap && c->n_real <= e->call_ld && !e->call_pool.empty()) {
      c->buf = e->call_pool.back();
      e->call_pool.pop_back();
    }
  }
  if (!c->buf.pcm) {
    CallBuf b;
    b.ld = (c->n_real + 15) / 16 * 16;
    b.pcm_cap = n_samples;
    BW_CUDA(cudaMalloc(reinterpret_cast<void**>(&b.pcm), (size_t)(n_samples + 4) * 4));
    BW_CUDA(cudaMalloc(reinterpret_cast<void**>(&b.logmel), (size_t)e->dims.n_mels * b.ld * 4));
    BW_CUDA(cudaMalloc(reinterpret_cast<void**>(&b.gmax), 4));
    c->buf = b;
  }
  BW_CUDA(cudaEventCreateWithFlags(&c->mel_done, cudaEventDisableTiming));
  compute_call_mel(e, c.get(), pcm, n_samples);
  *out = c.release();
  BW_API_END
}

int bw_engine_set_resampler(bw_engine* e, int32_t sample_rate, int32_t orig_freq, int32_t new_freq, int32_t width, const float* taps) {
  BW_API_BEGIN
  BW_CHECK(e && taps, "null argument");
  BW_CHECK(sample_rate > 0 && sample_rate != 16000, "16 kHz needs no resampler");
  BW_CHECK(orig_freq > 0 && new_freq > 0 && width > 0 && orig_freq <= 4096 && new_freq <= 4096 && width <= 4096, "bad filter geometry");
  BW_CHECK((long long)sample_rate * new_freq == 16000LL * orig_freq, "orig_freq / new_freq must equal sample_rate / 16000");
  DeviceGuard dg(e->device);
  auto r = std::make_unique<Resampler>();
  r->orig = orig_freq; r->nw = new_freq; r->width = width; r->K = 2 * width + orig_freq;
  std::vector<int2> ranges(new_freq);
  for (int i = 0; i < new_freq; ++i) {
    int lo = r->K, hi = 0;
    for (int k = 0; k < r->K; ++k)
      if (fabsf(taps[(size_t)i * r->K + k]) >= 1e-20f) { lo = std::min(lo, k); hi = std::max(hi, k + 1); }
    if (lo > hi) { lo = 0; hi = 0; }
    ranges[i] = make_int2(lo, hi);
  }
  r->taps.alloc((size_t)new_freq * r->K * 4);
  r->ranges.alloc((size_t)new_freq * sizeof(int2));
  BW_CUDA(cudaMemcpy(r->taps.p, taps, (size_t)new_freq * r->K * 4, cudaMemcpyHostToDevice));
  BW_CUDA(cudaMemcpy(r->ranges.p, ranges.data(), (size_t)new_freq * sizeof(int2), cudaMemcpyHostToDevice));
  std::lock_guard<std::mutex> g(e->resampler_mu);
  BW_CHECK(e->resamplers.find(sample_rate) == e->resamplers.end(), "resampler already registered for this rate");
  e->resamplers[sample_rate] = std::move(r);
  BW_API_END
}

int bw_call_open_pcm16(bw_engine* e, const int16_t* pcm, int64_t n_in, int32_t sample_rate, bw_call** out) {
  BW_API_BEGIN
  BW_CHECK(e && pcm && out, "null argument");
  BW_CHECK(e->state == 1, "engine not finalized");
  BW_CHECK(n_in >= 1, "empty audio");
  DeviceGuard dg(e->device);
  const Resampler* r = find_resampler(e, sample_rate);
  const long long n_samples = resampled_length(r, n_in);
  BW_CHECK(n_samples >= 1, "empty audio after resampling");
  auto c = std::make_unique<bw_call>();
  c->eng = e;
  c->n_samples = n_samples;
  c->total_frames = (int)((n_samples + 480000) / 160);
  c->content_frames = c->total_frames - 3000;
  c->n_real = (int)std::min<long long>(c->total_frames, (n_samples + 200 + 159) / 160);
  {
    std::lock_guard<std::mutex> g(e->call_mu);
    if (n_samples <= e->call_pcm_cap && c->n_real <= e->call_ld && !e->call_pool.empty()) {
      c->buf = e->call_pool.back();
      e->call_pool.pop_back();
    }
  }
  if (!c->buf.pcm) {
    CallBuf b;
    b.ld = (c->n_real + 15) / 16 * 16;
    b.pcm_cap = n_samples;
    BW_CUDA(cudaMalloc(reinterpret_cast<void**>(&b.pcm), (size_t)(n_samples + 4) * 4));
    BW_CUDA(cudaMalloc(reinterpret_cast<void**>(&b.logmel), (size_t)e->dims.n_mels * b.ld * 4));
    BW_CUDA(cudaMalloc(reinterpret_cast<void**>(&b.gmax), 4));
    c->buf = b;
  }
  BW_CUDA(cudaEventCreateWithFlags(&c->mel_done, cudaEventDisableTiming));
  {
    const unsigned idx = e->front_rr.fetch_add(1) % bw_engine::kFrontStreams;
    std::lock_guard<std::mutex> g(e->front_mu[idx]);
    cudaStream_t st = e->front[idx];
    ingest_pcm16(e, (int)idx, st, pcm, n_in, r, c->buf.pcm, n_samples);
    mel_power(c->buf.pcm, n_samples, 480000, e->mel_tables.as<float>(), e->mel_filters.as<float>(), e->mel_ranges.as<int2>(),
              e->dims.n_mels, c->buf.logmel, c->buf.ld, c->n_real, c->total_frames, c->buf.gmax, st);
    BW_CUDA(cudaEventRecord(c->mel_done, st));
  }
  *out = c.release();
  BW_API_END
}

int bw_resample_pcm16(bw_engine* e, const int16_t* pcm, int64_t n_in, int32_t sample_rate, float* out, int64_t* n_out) {
  BW_API_BEGIN
  BW_CHECK(e && pcm && out && n_out, "null argument");
  BW_CHECK(n_in >= 1, "empty audio");
  DeviceGuard dg(e->device);
  const Resampler* r = find_resampler(e, sample_rate);
  const long long n16 = resampled_length(r, n_in);
  DevBuf dst;
  dst.alloc((size_t)(n16 + 4) * 4);
  std::lock_guard<std::mutex> g(e->front_mu[0]);
  ingest_pcm16(e, 0, e->front[0], pcm, n_in, r, dst.as<float>(), n16);
  BW_CUDA(cudaMemcpyAsync(out, dst.p, (size_t)n16 * 4, cudaMemcpyDeviceToHost, e->front[0]));
  BW_CUDA(cudaStreamSynchronize(e->front[0]));
  *n_out = n16;
  BW_API_END
}

int bw_call_content_frames(bw_call* c, int32_t* out) {
  if (!c || !out) return BW_ERR_INVALID;
  *out = c->content_frames;
  return BW_OK;
}

int bw_call_close(bw_call* c) {
  BW_API_BEGIN
  if (!c) return BW_OK;
  bw_engine* e = c->eng;
  DeviceGuard dg(e->device);
  cudaEventSynchronize(c->mel_done);
  cudaEventDestroy(c->mel_done);
  if (c->buf.pooled) {
    std::lock_guard<std::mutex> g(e->call_mu);
    e->call_pool.push_back(c->buf);
  } else {
    cudaFree(c->buf.pcm); cudaFree(c->buf.logmel); cudaFree(c->buf.gmax);
  }
  delete c;
  BW_API_END
}

int bw_call_decode(bw_call* c, int32_t seek, const bw_decode_opts* o, bw_result* out) {
  BW_API_BEGIN
  BW_CHECK(c && o && out, "null argument");
  bw_engine* e = c->eng;
  BW_CHECK(o->n_initial >= 1 && o->initial_tokens, "initial tokens required");
  BW_CHECK(o->n_initial < e->dims.n_text_ctx, "too many initial tokens");
  BW_CHECK(o->n_initial <= 448, "too many initial tokens");
  BW_CHECK(o->sot_index >= 0 && o->sot_index < o->n_initial, "sot_index out of range");
  BW_CHECK(o->beam_size >= 0 && o->beam_size <= kMaxBeam, "beam_size must be in [0, 8]");
  BW_CHECK(seek >= 0, "negative seek");
  for (int i = 0; i < o->n_initial; ++i) BW_CHECK(o->initial_tokens[i] >= 0 && o->initial_tokens[i] < e->dims.n_vocab, "token id out of range");
  Request r;
  r.kind = REQ_DECODE;
  r.call = c;
  r.seek = seek;
  r.initial.assign(o->initial_tokens, o->initial_tokens + o->n_initial);
  r.sot_index = o->sot_index;
  r.greedy = o->beam_size == 0;
  r.beam = o->beam_size;
  r.G = r.greedy ? 1 : o->beam_size;
  if (o->temperature > 0.f) {
    // upstream DecodingTask: GreedyDecoder(temperature) with n_group = best_of or 1; beam search is a T = 0 decoder
    BW_CHECK(r.greedy, "temperature > 0 needs beam_size == 0 (decode_with_fallback drops beam_size / patience above T = 0)");
    BW_CHECK(o->best_of >= 0 && o->best_of <= kMaxBeam, "best_of must be in [0, 8]");
    BW_CHECK(std::isfinite(o->temperature), "temperature must be finite");
    r.temperature = o->temperature;
    r.G = std::max(1, (int)o->best_of);
    r.seed = ((unsigned long long)o->seed_hi << 32) | (unsigned long long)o->seed_lo;
  }
  r.patience = o->patience > 0 ? o->patience : 1.f;
  r.length_penalty = o->length_penalty;
  r.sample_len = o->sample_len > 0 ? o->sample_len : e->dims.n_text_ctx / 2;
  r.without_ts = o->without_timestamps != 0;
  r.suppress_blank = o->suppress_blank != 0;
  r.max_initial_ts = o->max_initial_timestamp_index;
  BW_CHECK(lround((double)r.G * r.patience) <= kMaxFinished, "beam_size * patience too large");
  memset(out, 0, sizeof(*out));
  r.out = out;
  const int st = submit_and_wait(e, r);
  if (st != BW_OK) return st;
  BW_API_END
}

int bw_call_detect_language(bw_call* c, int32_t seek, bw_lang_result* out) {
  BW_API_BEGIN
  BW_CHECK(c && out, "null argument");
  bw_engine* e = c->eng;
  BW_CHECK(e->tt.num_languages > 0, "model has no language tokens");
  Request r;
  r.kind = REQ_LANG;
  r.call = c;
  r.seek = seek;
  r.initial = {e->tt.sot};
  r.G = 1;
  memset(out, 0, sizeof(*out));
  r.lang_out = out;
  const int st = submit_and_wait(e, r);
  if (st != BW_OK) return st;
  BW_API_END
}

// ---- stage level ----
int bw_mel(bw_engine* e, const float* pcm, int64_t n, int32_t padding, float* out_mel, int32_t* out_frames) {
  BW_API_BEGIN
  BW_CHECK(e && pcm && out_mel && out_frames, "null argument");
  BW_CHECK(e->filters_set, "mel filters not set");
  BW_CHECK(n + padding > 400, "audio (+padding) must be longer than 400 samples");
  DeviceGuard dg(e->device);
  const int nm = e->dims.n_mels;
  const int total = (int)((n + padding) / 160);
  const int n_real = padding >= 200 ? (int)std::min<long long>(total, (n + 200 + 159) / 160) : total;
  const int ld = (std::max(n_real, 1) + 15) / 16 * 16;
  DevBuf pcm_d, logmel, gmax, outd;
  pcm_d.alloc((size_t)(n + 4) * 4);
  logmel.alloc((size_t)nm * ld * 4);
  gmax.alloc(4);
  outd.alloc((size_t)nm * std::max(total, 1) * 4);
  std::lock_guard<std::mutex> g(e->gpu_mu);
  BW_CUDA(cudaMemcpyAsync(pcm_d.p, pcm, (size_t)n * 4, cudaMemcpyHostToDevice, e->stream));
  mel_power(pcm_d.as<float>(), n, padding, e->mel_tables.as<float>(), e->mel_filters.as<float>(), e->mel_ranges.as<int2>(), nm,
            logmel.as<float>(), ld, n_real, total, gmax.as<int>(), e->stream);
  mel_normalize_f32(logmel.as<float>(), ld, n_real, gmax.as<int>(), nm, total, outd.as<float>(), e->stream);
  BW_CUDA(cudaMemcpyAsync(out_mel, outd.p, (size_t)nm * total * 4, cudaMemcpyDeviceToHost, e->stream));
  BW_CUDA(cudaStreamSynchronize(e->stream));
  *out_frames = total;
  BW_API_END
}

int bw_encode(bw_engine* e, const float* mel, int32_t batch, float* out) {
  BW_API_BEGIN
  BW_CHECK(e && mel && out, "null argument");
  BW_CHECK(e->state == 1, "engine not finalized");
  BW_CHECK(batch >= 1 && batch <= e->Be, "batch exceeds max_encoder_batch");
  DeviceGuard dg(e->device);
  const auto& d = e->dims;
  const size_t per = (size_t)d.n_mels * 3000;
  DevBuf tmp, outf;
  tmp.alloc(per * batch * 4);
  const size_t n_out = (size_t)batch * 1500 * d.n_audio_state;
  outf.alloc(n_out * 4);
  std::lock_guard<std::mutex> g(e->gpu_mu);
  BW_CUDA(cudaMemcpyAsync(tmp.p, mel, per * batch * 4, cudaMemcpyHostToDevice, e->stream));
  for (int b = 0; b < batch; ++b) engine_window_to_A1(e, tmp.as<float>() + per * b, 3000, 3000, nullptr, 0, 3000, b);
  engine_encoder_forward(e, batch);
  if (e->fp32) BW_CUDA(cudaMemcpyAsync(out, e->enc_out.p, n_out * 4, cudaMemcpyDeviceToHost, e->stream));
  else {
    f32_from_bf16(e->enc_out.as<bf16>(), outf.as<float>(), (long long)n_out, e->stream);
    BW_CUDA(cudaMemcpyAsync(out, outf.p, n_out * 4, cudaMemcpyDeviceToHost, e->stream));
  }
  BW_CUDA(cudaStreamSynchronize(e->stream));
  BW_API_END
}

int bw_decode_logits(bw_engine* e, const float* mel_window, const int32_t* tokens, int32_t n, float* out_logits) {
  BW_API_BEGIN
  BW_CHECK(e && mel_window && tokens && out_logits, "null argument");
  BW_CHECK(n >= 1 && n <= e->dims.n_text_ctx, "token count out of range");
  Request r;
  r.kind = REQ_LOGITS;
  r.host_mel = mel_window;
  r.initial.assign(tokens, tokens + n);
  r.G = 1;
  r.logits_out = out_logits;
  const int st = submit_and_wait(e, r);
  if (st != BW_OK) return st;
  BW_API_END
}

// ---- kernel-level entry points (device pointers / resident synthetic data; used by tests and bench.py) ----
int bw_gemm_bf16(int impl, const void* A, const void* B, void* C, const float* bias, const float* residual, int32_t M, int32_t N,
                 int32_t K, int32_t gelu, int32_t out_fp32, void* stream) {
  BW_API_BEGIN
  GemmArgs g;
  g.A = A; g.B = B; g.C = C; g.bias = bias; g.residual = residual;
  g.M = M; g.N = N; g.K = K; g.lda = K; g.ldb = K; g.ldc = N; g.ldres = N; g.gelu = gelu != 0; g.out_fp32 = out_fp32 != 0;
  if (impl == 0) gemm_tc_bf16(g, reinterpret_cast<cudaStream_t>(stream));
  else if (impl == 2) {  // swap-AB path of the decoder: C[M, N] computed as (B . A^T)^T
    GemmArgs s = g;
    s.A = B; s.B = A; s.M = N; s.N = M; s.transposed = true;
    gemm_tc_bf16(s, reinterpret_cast<cudaStream_t>(stream));
  } else gemm_simt<bf16>(g, reinterpret_cast<cudaStream_t>(stream));
  BW_API_END
}

// Test hook for the decoder LayerNorm fusion: a producer row GEMM (x = res + A.Wp^T + bp, which also leaves bf16(x)
// and the LayerNorm partials) followed by a consumer row GEMM (out = [gelu](LayerNorm(x).Wc^T + bc) with the
// LayerNorm folded into Wc).  If A is null the producer is skipped and x = res goes through rows_ln_partials (the
// embedding path).  All pointers are device pointers; x_out fp32 [M, d], out fp32 [M, N].
int bw_test_ln_chain(const void* A, const void* Wp, const float* bp, const float* res, const float* gamma, const float* beta,
                     const float* Wc, const float* bc, int32_t M, int32_t d, int32_t Kp, int32_t N, int32_t gelu, float* x_out,
                     float* out, void* stream) {
  BW_API_BEGIN
  BW_CHECK(res && gamma && beta && Wc && x_out && out && M > 0 && d > 0 && N > 0 && d % 64 == 0 && N % 64 == 0, "bad argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  DevBuf Wf, c1, c2, xb, lst;
  Wf.alloc((size_t)N * d * 2); c1.alloc((size_t)N * 4); c2.alloc((size_t)N * 4);
  xb.alloc((size_t)M * d * 2); lst.alloc((size_t)M * (d / 64) * sizeof(float2));
  fold_layernorm(Wc, gamma, beta, bc, N, d, Wf.as<bf16>(), c1.as<float>(), c2.as<float>(), st);
  BW_CUDA(cudaMemcpyAsync(x_out, res, (size_t)M * d * 4, cudaMemcpyDeviceToDevice, st));
  if (A) {
    GemmArgs g;
    g.A = A; g.B = Wp; g.M = M; g.N = d; g.K = Kp; g.lda = Kp; g.ldb = Kp; g.ldc = d; g.ldres = d;
    g.bias = bp; g.residual = x_out; g.C = x_out; g.out_fp32 = true; g.xb_out = xb.p; g.ln_stats_out = lst.as<float2>();
    gemm_tc_rows(g, st);
  } else {
    rows_ln_partials(x_out, M, d, xb.as<bf16>(), lst.as<float2>(), st);
  }
  GemmArgs c;
  c.A = xb.p; c.B = Wf.p; c.M = M; c.N = N; c.K = d; c.lda = d; c.ldb = d; c.ldc = N;
  c.bias = c2.as<float>(); c.C = out; c.out_fp32 = true; c.gelu = gelu != 0;
  c.ln_stats_in = lst.as<float2>(); c.ln_c1 = c1.as<float>();
  gemm_tc_rows(c, st);
  BW_CUDA(cudaStreamSynchronize(st));
  BW_API_END
}

int bw_attention_bf16(int impl, const void* qkv, void* out, int32_t batch, int32_t T_len, int32_t n_head, void* stream) {
  BW_API_BEGIN
  if (impl == 0) attn_encoder_tc(reinterpret_cast<const bf16*>(qkv), reinterpret_cast<bf16*>(out), batch, T_len, n_head, reinterpret_cast<cudaStream_t>(stream));
  else attn_encoder_simt<bf16>(reinterpret_cast<const bf16*>(qkv), reinterpret_cast<bf16*>(out), batch, T_len, n_head, reinterpret_cast<cudaStream_t>(stream));
  BW_API_END
}

namespace {
struct EvTimer {
  cudaEvent_t a = nullptr, b = nullptr;
  cudaStream_t st;
  explicit EvTimer(cudaStream_t s) : st(s) { BW_CUDA(cudaEventCreate(&a)); BW_CUDA(cudaEventCreate(&b)); }
  ~EvTimer() { cudaEventDestroy(a); cudaEventDestroy(b); }
  void start() { BW_CUDA(cudaEventRecord(a, st)); }
  float stop_ms() { BW_CUDA(cudaEventRecord(b, st)); BW_CUDA(cudaEventSynchronize(b)); float ms = 0; BW_CUDA(cudaEventElapsedTime(&ms, a, b)); return ms; }
};
}  // namespace

int bw_bench_mel(bw_engine* e, int64_t n, int32_t iters, float* ms_out, double* bytes_out) {
  BW_API_BEGIN
  BW_CHECK(e && ms_out && bytes_out && iters > 0 && n > 400, "bad argument");
  DeviceGuard dg(e->device);
  const int nm = e->dims.n_mels;
  const int total = (int)((n + 480000) / 160);
  const int n_real = (int)std::min<long long>(total, (n + 200 + 159) / 160);
  const int ld = (n_real + 15) / 16 * 16;
  DevBuf pcm, logmel, gmax, a1;
  pcm.alloc((size_t)(n + 4) * 4); logmel.alloc((size_t)nm * ld * 4); gmax.alloc(4);
  BW_CUDA(cudaMemset(pcm.p, 0, pcm.bytes));
  std::lock_guard<std::mutex> g(e->gpu_mu);
  EvTimer t(e->stream);
  auto once = [&] {
    mel_power(pcm.as<float>(), n, 480000, e->mel_tables.as<float>(), e->mel_filters.as<float>(), e->mel_ranges.as<int2>(), nm,
              logmel.as<float>(), ld, n_real, total, gmax.as<int>(), e->stream);
    engine_window_to_A1(e, logmel.as<float>(), ld, n_real, gmax.as<int>(), 0, std::min(3000, total - 3000), 0);
  };
  once();
  t.start();
  for (int i = 0; i < iters; ++i) once();
  *ms_out = t.stop_ms() / iters;
  // 4*(n+padding) read + 4*n_mels*(n+padding)/160 written
  *bytes_out = 4.0 * (double)(n + 480000) + 4.0 * nm * (double)total;
  BW_API_END
}

int bw_bench_encoder(bw_engine* e, int32_t batch, int32_t iters, float* ms_out, double* flops_out) {
  BW_API_BEGIN
  BW_CHECK(e && ms_out && flops_out && iters > 0, "bad argument");
  BW_CHECK(e->state == 1 && batch >= 1 && batch <= e->Be, "batch exceeds max_encoder_batch");
  DeviceGuard dg(e->device);
  std::lock_guard<std::mutex> g(e->gpu_mu);
  BW_CUDA(cudaMemsetAsync(e->A1.p, 0, e->A1.bytes, e->stream));
  engine_encoder_forward(e, batch);
  EvTimer t(e->stream);
  t.start();
  for (int i = 0; i < iters; ++i) engine_encoder_forward(e, batch);
  *ms_out = t.stop_ms() / iters;
  const double d = e->dims.n_audio_state, L = e->dims.n_audio_layer, nm = e->dims.n_mels;
  *flops_out = batch * (2.0 * 3000 * 3 * nm * d + 2.0 * 1500 * 3 * d * d + L * (8.0 * 1500 * d * d + 4.0 * 1500 * 1500 * d + 16.0 * 1500 * d * d));
  BW_API_END
}

namespace {
// Synthetic resident decode: `n_segments` windows x `n_group` hypotheses, positions [start_len, start_len + n_steps).
// Same launches, control upload and per-step completion read-back as the scheduler's decode_step().
void synthetic_init(bw_engine* e, int n_segments, int n_group, int start_len) {
  for (int i = 0; i < n_segments; ++i) {
    int* rec = e->h_init + i * kInitRecInts;
    rec[0] = i; rec[1] = n_group; rec[2] = 0; rec[3] = 3; rec[4] = start_len; rec[5] = i * n_group; rec[6] = 0; rec[7] = 1;
    rec[8] = 50; rec[9] = kMaxFinished; rec[10] = e->tt.timestamp_begin - 1000;
    rec[11] = rec[12] = rec[13] = rec[14] = rec[15] = 0;
  }
  BW_CUDA(cudaMemcpyAsync(e->d_init.p, e->h_init, (size_t)n_segments * kInitRecInts * 4, cudaMemcpyHostToDevice, e->stream));
  engine_init_requests(e, e->d_init.as<int>(), n_segments);
  BW_CUDA(cudaStreamSynchronize(e->stream));
}
// same grouping, launches, control upload and per-step completion read-back as the scheduler's decode_step()
void synthetic_step(bw_engine* e, Ctl* ctls, int n_segments, int n_group, int cur) {
  const int ng = choose_groups(n_segments);
  for (int g = 0; g < ng; ++g) ctls[g].reset();
  for (int i = 0; i < n_segments; ++i) {
    Ctl& c = ctls[i % ng];
    c.grp_first[c.NG] = c.R; c.grp_n[c.NG] = n_group; c.grp_x[c.NG] = i; ++c.NG;
    c.max_grp = std::max(c.max_grp, n_group);
    c.act_req[c.NA] = i; c.act_first[c.NA] = c.SR; ++c.NA;
    for (int j = 0; j < n_group; ++j) {
      c.row_seq[c.R] = i * n_group + j; c.row_pos[c.R] = cur - 1; c.row_tok[c.R] = -1; c.row_bpos[c.R] = cur - 1;
      c.lrow_src[c.LR] = c.R; c.srow_lrow[c.SR] = c.LR; c.srow_req[c.SR] = i; c.srow_seq[c.SR] = i * n_group + j;
      ++c.R; ++c.LR; ++c.SR;
    }
  }
  for (int g = 0; g < ng; ++g) enqueue_group_step(e, e->grp[g], ctls[g]);
  for (int g = 0; g < ng; ++g) BW_CUDA(cudaStreamSynchronize(e->grp[g].stream));
  BW_CUDA(cudaMemcpyAsync(e->h_flags, e->rs.completed, (size_t)e->Q * 4, cudaMemcpyDeviceToHost, e->stream));
  BW_CUDA(cudaStreamSynchronize(e->stream));
  e->anc_cur ^= 1;
}
struct SynCtls {
  Ctl c[kMaxGroups];
  explicit SynCtls(bw_engine* e) { for (int g = 0; g < kMaxGroups; ++g) c[g].layout(e->grp[g].h_ctrl, e->R_max, e->LR_max, e->Q); }
};
}  // namespace

// One full decoder step (all layers + logits + sampling + beam update) over `n_segments` resident
// windows with `n_group` hypotheses each at context length `context_len`, timed with CUDA events.
int bw_bench_decoder_step(bw_engine* e, int32_t n_segments, int32_t n_group, int32_t context_len, int32_t iters, float* ms_out,
                          double* bytes_out) {
  BW_API_BEGIN
  BW_CHECK(e && ms_out && bytes_out && iters > 0, "bad argument");
  BW_CHECK(e->state == 1, "engine not finalized");
  BW_CHECK(n_segments >= 1 && n_segments <= e->Q && n_group >= 1 && n_group <= kMaxBeam && n_segments * n_group <= e->S, "exceeds pools");
  BW_CHECK(context_len >= 4 && context_len + iters + 2 < e->dims.n_text_ctx, "context_len out of range");
  DeviceGuard dg(e->device);
  std::lock_guard<std::mutex> g(e->gpu_mu);
  BW_CHECK(e->live.empty(), "engine busy");
  SynCtls sc(e);
  BW_CUDA(cudaMemsetAsync(e->cross_cache.p, 0, (size_t)n_segments * (e->cross_cache.bytes / e->Q), e->stream));
  synthetic_init(e, n_segments, n_group, context_len);
  int cur = context_len;
  synthetic_step(e, sc.c, n_segments, n_group, cur++);
  EvTimer t(e->stream);  // e->stream is idle here and receives the completion read-back of every step
  t.start();
  for (int i = 0; i < iters; ++i) synthetic_step(e, sc.c, n_segments, n_group, cur++);
  *ms_out = t.stop_ms() / iters;
  const double ts = e->fp32 ? 4 : 2, d = e->dims.n_text_state, L = e->dims.n_text_layer, V = e->dims.n_vocab;
  const double S = (double)n_segments * n_group;
  *bytes_out = ts * (L * 14 * d * d + V * d) + n_segments * ts * L * 2 * 1500 * d + S * ts * L * 2 * (context_len + iters / 2.0) * d + 4 * S * V;
  BW_API_END
}

// The whole hot path on device-resident PCM: log-mel -> encoder (batches of max_encoder_batch) -> cross-KV ->
// `n_steps` batched decoder steps for `n_segments` windows of `n_samples` samples each.  One CUDA-event pair
// on the engine stream brackets everything (bench.py `value`: inputs resident in HBM when timing starts).
int bw_bench_pipeline(bw_engine* e, const float* pcm_host, const int64_t* offsets, const int64_t* lengths, int32_t n_segments,
                      int32_t n_group, int32_t n_steps, float* ms_out) {
  BW_API_BEGIN
  BW_CHECK(e && pcm_host && offsets && lengths && ms_out, "bad argument");
  BW_CHECK(e->state == 1, "engine not finalized");
  BW_CHECK(n_segments >= 1 && n_segments <= e->Q && n_group >= 1 && n_group <= kMaxBeam && n_segments * n_group <= e->S, "exceeds pools");
  BW_CHECK(n_steps >= 1 && 3 + n_steps < e->dims.n_text_ctx, "n_steps out of range");
  for (int i = 0; i < n_segments; ++i) BW_CHECK(lengths[i] > 400 && lengths[i] <= e->call_pcm_cap, "segment length out of range");
  DeviceGuard dg(e->device);
  std::lock_guard<std::mutex> g(e->gpu_mu);
  BW_CHECK(e->live.empty(), "engine busy");
  std::vector<CallBuf> bufs;
  {
    std::lock_guard<std::mutex> cg(e->call_mu);
    BW_CHECK((int)e->call_pool.size() >= n_segments, "not enough call buffers");
    for (int i = 0; i < n_segments; ++i) { bufs.push_back(e->call_pool.back()); e->call_pool.pop_back(); }
  }
  for (int i = 0; i < n_segments; ++i)
    BW_CUDA(cudaMemcpyAsync(bufs[i].pcm, pcm_host + offsets[i], (size_t)lengths[i] * 4, cudaMemcpyHostToDevice, e->stream));
  BW_CUDA(cudaStreamSynchronize(e->stream));
  SynCtls sc(e);
  auto frames = [&](int i, int& total, int& n_real, int& seg) {
    total = (int)((lengths[i] + 480000) / 160);
    n_real = (int)std::min<long long>(total, (lengths[i] + 200 + 159) / 160);
    seg = std::min(3000, total - 3000);
  };
  EvTimer t(e->stream);
  t.start();
  for (int i = 0; i < n_segments; ++i) {
    int total, n_real, seg;
    frames(i, total, n_real, seg);
    mel_power(bufs[i].pcm, lengths[i], 480000, e->mel_tables.as<float>(), e->mel_filters.as<float>(), e->mel_ranges.as<int2>(),
              e->dims.n_mels, bufs[i].logmel, bufs[i].ld, n_real, total, bufs[i].gmax, e->stream);
  }
  for (int s0 = 0; s0 < n_segments; s0 += e->Be) {
    const int nb = std::min(e->Be, n_segments - s0);
    for (int i = 0; i < nb; ++i) {
      int total, n_real, seg;
      frames(s0 + i, total, n_real, seg);
      engine_window_to_A1(e, bufs[s0 + i].logmel, bufs[s0 + i].ld, n_real, bufs[s0 + i].gmax, 0, seg, i);
    }
    engine_encoder_forward(e, nb);
    for (int i = 0; i < nb; ++i) engine_cross_kv(e, i, s0 + i);
  }
  synthetic_init(e, n_segments, n_group, 3);
  for (int i = 0; i < n_steps; ++i) synthetic_step(e, sc.c, n_segments, n_group, 3 + i);
  *ms_out = t.stop_ms();
  {
    std::lock_guard<std::mutex> cg(e->call_mu);
    for (auto& b : bufs) e->call_pool.push_back(b);
  }
  BW_API_END
}

// The decoder step's dominant kernel alone: cross-attention of one layer over resident K/V.
int bw_bench_cross_attention(bw_engine* e, int32_t n_segments, int32_t n_group, int32_t iters, float* ms_out, double* bytes_out) {
  BW_API_BEGIN
  BW_CHECK(e && ms_out && bytes_out && iters > 0, "bad argument");
  BW_CHECK(e->state == 1 && n_segments >= 1 && n_segments <= e->Q && n_group >= 1 && n_group <= kMaxBeam, "exceeds pools");
  BW_CHECK(n_segments * n_group <= e->R_max, "too many rows");
  DeviceGuard dg(e->device);
  std::lock_guard<std::mutex> g(e->gpu_mu);
  SynCtls sc(e);
  Ctl& ctl = sc.c[0];
  DecGroup& G = e->grp[0];
  for (int i = 0; i < n_segments; ++i) { ctl.grp_first[i] = i * n_group; ctl.grp_n[i] = n_group; ctl.grp_x[i] = i; }
  int* dbase = G.d_ctrl.as<int>();
  auto dev = [&](int* h) { return dbase + (h - ctl.base); };
  BW_CUDA(cudaMemcpyAsync(dbase, ctl.base, ctl.total * 4, cudaMemcpyHostToDevice, e->stream));
  BW_CUDA(cudaMemsetAsync(e->cross_cache.p, 0, (size_t)n_segments * (e->cross_cache.bytes / e->Q), e->stream));
  const auto& d = e->dims;
  const int dm = d.n_text_state, L = d.n_text_layer;
  const int R = n_segments * n_group;
  auto run = [&](int layer) {
    CrossKV x; x.cache = e->cross_cache.p; x.slot_stride = (long long)L * d.n_audio_ctx * 2 * dm; x.T_enc = d.n_audio_ctx;
    x.n_slots = e->Q; x.n_layer = L;
    if (e->fp32)
      dec_cross_attention<float>(dev(ctl.grp_first), dev(ctl.grp_n), dev(ctl.grp_x), n_segments, n_group, R, G.d_q.as<float>(), x, layer, dm,
                                 d.n_text_head, G.d_att.as<float>(), G.d_ws.as<float>(), e->stream);
    else
      dec_cross_attention<bf16>(dev(ctl.grp_first), dev(ctl.grp_n), dev(ctl.grp_x), n_segments, n_group, R, G.d_q.as<float>(), x, layer, dm,
                                d.n_text_head, G.d_att.as<bf16>(), G.d_ws.as<float>(), e->stream);
  };
  run(0);
  EvTimer t(e->stream);
  t.start();
  // walk the layers so that consecutive launches touch different K/V (as in the real step): inputs > L2
  for (int i = 0; i < iters; ++i) run(i % L);
  *ms_out = t.stop_ms() / iters;
  const double ts = e->fp32 ? 4 : 2;
  *bytes_out = (double)n_segments * ts * 2 * 1500 * dm + (double)R * dm * ts * 2;
  BW_API_END
}

int bw_debug_trace(bw_engine* e, int32_t enable, uint64_t* out, int32_t cap, int32_t* n_out) {
  BW_API_BEGIN
  // e may be null (kernel-level tools): then the current device is used
  int cur_dev = 0;
  BW_CUDA(cudaGetDevice(&cur_dev));
  DeviceGuard dg(e ? e->device : cur_dev);
  static std::mutex no_engine_mu;
  std::lock_guard<std::mutex> g(e ? e->gpu_mu : no_engine_mu);
  BW_CUDA(cudaDeviceSynchronize());
  static DevBuf buf;
  const size_t bytes = (1 + 2 * (size_t)kTraceCap) * 8;
  if (enable == 1) {
    if (buf.bytes < bytes) buf.alloc(bytes);
    BW_CUDA(cudaMemset(buf.p, 0, bytes));
    g_trace_dev = buf.as<unsigned long long>();
  } else if (enable == 2) {  // raw dump of the first `cap` records' worth of the buffer (fixed-slot users), then disarm
    BW_CHECK(out && cap >= 0, "null argument");
    unsigned long long* dev = g_trace_dev;
    g_trace_dev = nullptr;
    if (dev) BW_CUDA(cudaMemcpy(out, dev, std::min(bytes, (size_t)cap * 16), cudaMemcpyDeviceToHost));
  } else {
    BW_CHECK(out && n_out && cap >= 0, "null argument");
    unsigned long long* dev = g_trace_dev;
    g_trace_dev = nullptr;
    *n_out = 0;
    if (dev) {
      unsigned long long cnt = 0;
      BW_CUDA(cudaMemcpy(&cnt, dev, 8, cudaMemcpyDeviceToHost));
      const int n = (int)std::min<unsigned long long>(std::min<unsigned long long>(cnt, kTraceCap), (unsigned long long)cap);
      if (n > 0) BW_CUDA(cudaMemcpy(out, dev + 1, (size_t)n * 16, cudaMemcpyDeviceToHost));
      *n_out = n;
    }
  }
  BW_API_END
}

}  // extern "C"
