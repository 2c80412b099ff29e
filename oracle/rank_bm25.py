"""ORACLE — test infrastructure, not product code.

CPU restatement of the third-party package ``rank-bm25==0.2.2`` (pinned in the
reference's requirements.txt:38; NOT vendored under /root/reference and NOT
installable here — no network).  The reference imports it at
src/rag/bm25_index.py:17 and calls it at :126, :153, :236, :265.

PARITY UNPINNED: the reference has no test, fixture or golden vector that pins
BM25Okapi outputs, and the real wheel is not reachable from this container, so
this file restates the published Okapi-BM25 algorithm of that release from
its documented behaviour:

  * per-document term-frequency dicts, ``doc_len``, ``avgdl = sum(len)/N``
  * ``idf[w] = ln(N - df + 0.5) - ln(df + 0.5)`` (natural log)
  * ``average_idf = sum(idf) / len(idf)`` summed in vocabulary first-seen order
  * every idf < 0 is replaced by ``epsilon * average_idf`` (epsilon = 0.25),
    using the average computed BEFORE the replacement
  * ``get_scores``: fp64 zeros(N); for each query token in order (duplicates
    are added again): ``score += (idf.get(q) or 0) * (tf*(k1+1) /
    (tf + k1*(1 - b + b*doc_len/avgdl)))`` with k1 = 1.5, b = 0.75

This module is importable under the name ``rank_bm25`` (the harness in
oracle/ref_harness.py puts it on sys.modules) so that the reference's own
src/rag/bm25_index.py runs unmodified around it.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` leg may import anything under oracle/.
"""
import math

import numpy as np

__all__ = ["BM25Okapi"]


class BM25Okapi:
    def __init__(self, corpus, tokenizer=None, k1=1.5, b=0.75, epsilon=0.25):
        self.k1 = k1
        self.b = b
        self.epsilon = epsilon
        self.tokenizer = tokenizer
        self.corpus_size = 0
        self.avgdl = 0
        self.doc_freqs = []
        self.idf = {}
        self.doc_len = []
        if tokenizer is not None:
            corpus = [tokenizer(doc) for doc in corpus]
        df = self._count(corpus)
        self._calc_idf(df)

    # term statistics ------------------------------------------------------
    def _count(self, corpus):
        df = {}          # word -> number of documents containing it (first-seen order)
        total_len = 0
        for doc in corpus:
            self.doc_len.append(len(doc))
            total_len += len(doc)
            tf = {}
            for w in doc:
                tf[w] = tf.get(w, 0) + 1
            self.doc_freqs.append(tf)
            for w in tf:
                df[w] = df.get(w, 0) + 1
            self.corpus_size += 1
        self.avgdl = total_len / self.corpus_size
        return df

    def _calc_idf(self, df):
        idf_sum = 0
        negative = []
        for w, n_w in df.items():
            v = math.log(self.corpus_size - n_w + 0.5) - math.log(n_w + 0.5)
            self.idf[w] = v
            idf_sum += v
            if v < 0:
                negative.append(w)
        self.average_idf = idf_sum / len(self.idf)
        floor = self.epsilon * self.average_idf
        for w in negative:
            self.idf[w] = floor

    # scoring --------------------------------------------------------------
    def get_scores(self, query):
        score = np.zeros(self.corpus_size)
        doc_len = np.array(self.doc_len)
        for q in query:
            q_freq = np.array([(d.get(q) or 0) for d in self.doc_freqs])
            score += (self.idf.get(q) or 0) * (
                q_freq * (self.k1 + 1)
                / (q_freq + self.k1 * (1 - self.b + self.b * doc_len / self.avgdl))
            )
        return score

    def get_batch_scores(self, query, doc_ids):
        assert all(di < len(self.doc_freqs) for di in doc_ids)
        score = np.zeros(len(doc_ids))
        doc_len = np.array(self.doc_len)[doc_ids]
        for q in query:
            q_freq = np.array([(self.doc_freqs[di].get(q) or 0) for di in doc_ids])
            score += (self.idf.get(q) or 0) * (
                q_freq * (self.k1 + 1)
                / (q_freq + self.k1 * (1 - self.b + self.b * doc_len / self.avgdl))
            )
        return score.tolist()

    def get_top_n(self, query, documents, n=5):
        assert self.corpus_size == len(documents)
        scores = self.get_scores(query)
        top = np.argsort(scores)[::-1][:n]
        return [documents[i] for i in top]
