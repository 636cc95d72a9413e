//! src/adapters/search/index_builder.rs — posts -> vocabulary (lexicographic term ids) + CSR.
//! Tested twins: openintel_b200/host/openintel_host.hpp (IndexBuilder), openintel_b200/store.py.
use std::collections::HashMap;

/// The reference tokenizer (src/adapters/analyzer/lexicon.rs:54-58): Unicode-lowercase, split on every
/// char that is not ASCII alphanumeric, drop empties.  No stemming, no stop words (src/domain/dip.rs:261-272).
pub fn tokenize(text: &str) -> Vec<String> {
    text.to_lowercase().split(|c: char| !c.is_ascii_alphanumeric()).filter(|w| !w.is_empty()).map(str::to_owned).collect()
}

#[derive(Default)]
pub struct IndexBuilder {
    pub(crate) vocab: Vec<String>,
    pub(crate) term_offsets: Vec<u64>,
    pub(crate) doc_ids: Vec<u32>,
    pub(crate) tfs: Vec<u32>,
    pub(crate) doc_len: Vec<u32>,
    pub(crate) post_ids: Vec<String>, // doc_id -> SocialPost::id (src/domain/entities/social_post.rs:31)
    raw: Vec<(String, u32, u32)>,     // (token, doc, tf), docs ascending
}

impl IndexBuilder {
    /// Documents get dense ids in input order.
    pub fn add(&mut self, post_id: &str, text: &str) {
        let doc = self.doc_len.len() as u32;
        let toks = tokenize(text);
        self.doc_len.push(toks.len() as u32);
        self.post_ids.push(post_id.to_owned());
        let mut tf: HashMap<String, u32> = HashMap::new();
        for t in toks {
            *tf.entry(t).or_insert(0) += 1;
        }
        for (t, f) in tf {
            self.raw.push((t, doc, f));
        }
    }
    pub fn finish(&mut self) {
        self.raw.sort_by(|a, b| a.0.cmp(&b.0).then(a.1.cmp(&b.1)));
        self.vocab.clear();
        self.term_offsets.clear();
        self.doc_ids.clear();
        self.tfs.clear();
        for (t, d, f) in &self.raw {
            if self.vocab.last() != Some(t) {
                self.vocab.push(t.clone());
                self.term_offsets.push(self.doc_ids.len() as u64);
            }
            self.doc_ids.push(*d);
            self.tfs.push(*f);
        }
        self.term_offsets.push(self.doc_ids.len() as u64);
    }
    pub fn n_docs(&self) -> usize {
        self.doc_len.len()
    }
    pub fn post_id(&self, doc_id: u32) -> Option<&str> {
        self.post_ids.get(doc_id as usize).map(String::as_str)
    }
    pub fn term_id(&self, token: &str) -> Option<u32> {
        self.vocab.binary_search_by(|v| v.as_str().cmp(token)).ok().map(|i| i as u32)
    }
    /// Known tokens of a query text as term ids, de-duplicated in first-seen order (the 64-term limit of
    /// docs/SPEC.md §3 counts DISTINCT terms; the library de-duplicates too, this just keeps the arrays small).
    pub fn query_terms(&self, text: &str) -> Vec<u32> {
        let mut out: Vec<u32> = Vec::new();
        for t in tokenize(text) {
            if let Some(id) = self.term_id(&t) {
                if !out.contains(&id) {
                    out.push(id);
                }
            }
        }
        out
    }
}
