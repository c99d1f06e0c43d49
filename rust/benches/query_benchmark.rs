//! Criterion query benchmark for the GENUINE probly-search crate (the reference ships only an
//! indexing bench, benches/test_benchmark.rs).  NOT RUN in this repository's environment (no
//! cargo); provided so anyone with a Rust toolchain can time the real CPU path on the cfg-0 shape:
//! 50k docs, 1 field, two Zipf-drawn tokens per doc, 1k single-term BM25 queries.
//!
//!   [dev-dependencies] criterion = "0.3", probly-search = "2.0.1"
use criterion::{criterion_group, criterion_main, Criterion};
use probly_search::{score::bm25, Index};
use std::borrow::Cow;

struct Doc { id: usize, title: String }
fn tokenizer(s: &str) -> Vec<Cow<'_, str>> { s.split(' ').map(Cow::from).collect() }
fn title(d: &Doc) -> Vec<&str> { vec![d.title.as_str()] }

struct SplitMix64(u64);
impl SplitMix64 {
    fn next(&mut self) -> u64 {
        self.0 = self.0.wrapping_add(0x9E3779B97F4A7C15);
        let mut z = self.0;
        z = (z ^ (z >> 30)).wrapping_mul(0xBF58476D1CE4E5B9);
        z = (z ^ (z >> 27)).wrapping_mul(0x94D049BB133111EB);
        z ^ (z >> 31)
    }
    fn uniform(&mut self) -> f64 { (self.next() >> 11) as f64 / 9007199254740992.0 }
}

fn bench(c: &mut Criterion) {
    // vocabulary + Zipf CDF exactly as probly_search_b200/csrc/workload.cpp (seed 0x5EEDC0DE + 0)
    let v = 1usize << 16;
    let mut r = SplitMix64(0x5EEDC0DE ^ 0x766F636162);
    let mut seen = std::collections::HashSet::new();
    let mut vocab = Vec::with_capacity(v);
    while vocab.len() < v {
        let len = 3 + (r.uniform() * 8.0) as usize;
        let s: String = (0..len).map(|_| (b'a' + (r.uniform() * 26.0) as u8) as char).collect();
        if seen.insert(s.clone()) { vocab.push(s); }
    }
    let h: f64 = (1..=v).map(|i| 1.0 / i as f64).sum();
    let mut cdf = Vec::with_capacity(v);
    let mut acc = 0.0;
    for i in 1..=v { acc += (1.0 / i as f64) / h; cdf.push(acc); }
    let zipf = |r: &mut SplitMix64| { let u = r.uniform(); cdf.partition_point(|&x| x <= u).min(v - 1) };

    let mut index = Index::<usize>::new_with_capacity(1, 100_000, 100_000);
    for d in 0..50_000usize {
        let mut rr = SplitMix64(0x5EEDC0DEu64.wrapping_add(0x9E3779B97F4A7C15u64.wrapping_mul(d as u64 + 1)));
        rr.next();
        let _n = rr.uniform();
        let title_s = format!("{} {}", vocab[zipf(&mut rr)], vocab[zipf(&mut rr)]);
        index.add_document(&[title], tokenizer, d, &Doc { id: d, title: title_s });
    }
    let mut qr = SplitMix64(0x9E3779B9);
    let queries: Vec<String> = (0..1000).map(|_| vocab[zipf(&mut qr)].clone()).collect();
    c.bench_function("query_1k_bm25_50k_docs", |b| {
        b.iter(|| { for q in &queries { criterion::black_box(index.query(q, &mut bm25::new(), tokenizer, &[1.0])); } })
    });
}
criterion_group!(benches, bench);
criterion_main!(benches);
