/* kc_host.h — host-side helpers around the engine: FASTA staging and the host tree.  Plain C ABI, no
 * CUDA types.  (The synthetic protein-set generator of the benchmarks lives in kc_synth.h / libkc_synth.so.)
 *
 * kc_fasta_* restates what the reference gets from seq_io (src/main.rs:62-72) and from
 * Protein::new / get_amr_class (src/protein.rs:107-110,135-138):
 *   id       = header text after '>' up to the first space or tab
 *   sequence = the record's residue bytes with line breaks removed
 *   class    = 4th '|'-separated field of the id (split_terminator semantics); records whose
 *              id has fewer than 4 fields get the empty class name (the reference would
 *              panic on them at src/protein.rs:137) and are counted in n_missing_class
 * ALL records are read regardless of file size, in file order.
 */
#ifndef KC_HOST_H_
#define KC_HOST_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kc_fasta kc_fasta;

/* threads = worker threads for parsing (the CLI's second argument, src/main.rs:59-60) */
int kc_fasta_parse_file(const char* path, int threads, kc_fasta** out);
int kc_fasta_parse_buffer(const char* data, uint64_t len, int threads, kc_fasta** out);
void kc_fasta_free(kc_fasta* f);
uint64_t kc_fasta_n_proteins(const kc_fasta* f);
uint64_t kc_fasta_n_residues(const kc_fasta* f);
const uint8_t* kc_fasta_residues(const kc_fasta* f);  /* page-locked when CUDA is present */
const uint64_t* kc_fasta_offsets(const kc_fasta* f);  /* n_proteins + 1 */
const uint32_t* kc_fasta_class_ids(const kc_fasta* f);
uint32_t kc_fasta_n_classes(const kc_fasta* f);
uint64_t kc_fasta_n_missing_class(const kc_fasta* f);
const char* kc_fasta_class_name(const kc_fasta* f, uint32_t class_id);
const char* kc_fasta_id(const kc_fasta* f, uint64_t protein);

/* DIAMOND hand-off of align_and_output_pairs (src/graph/mod.rs:195-319) without the `diamond` subprocesses:
 * creates <dir>/fasta_files and <dir>/db_files (:202-220), writes for edge i the two one-record files
 * fasta_files/{i}_{accession}.fasta with ">{id}\n{sequence}" (:253-261 reference = edges[i].a, :273-280 query =
 * edges[i].b; accession = id up to the first '|'), and starts <dir>/blastp_output.tsv with the reference's
 * header line (:304).  A DIAMOND step (makedb on the first file, blastp --outfmt 6 qseqid qlen sseqid slen
 * qstart qend sstart send length pident evalue bitscore on the second, :266-293) appends its rows to that
 * file.  The edge key is the index in the emitted list (the reference's is the racy index in its edge vector).
 * Needs kc_b200.h for kc_edge. */
struct kc_edge;
int kc_write_handoff(const kc_fasta* f, const struct kc_edge* edges, uint64_t n_edges, const char* dir,
                     uint64_t* n_files_out);
const char* kc_blastp_header(void);

/* Host tree clustering: the reference's src/tree.rs (Tree::new + add_protein for every protein in
 * input order, src/tree.rs:519-536) over the engine's per-protein id lists
 * (kc_get_protein_ids: row_offsets[n+1], ids ascending within a row, ids < n_ids).
 * Stays on the host (north star); see csrc/tree.cpp for what is cached. */
typedef struct kc_tree kc_tree;
int kc_tree_build(const uint64_t* row_offsets, const uint32_t* ids, uint64_t n_proteins, uint32_t n_ids,
                  kc_tree** out);
void kc_tree_free(kc_tree* t);
uint64_t kc_tree_n_merges(const kc_tree* t);     /* "Merging" events, src/tree.rs:227 */
uint64_t kc_tree_n_no_common(const kc_tree* t);  /* "No kmers in common" events, src/tree.rs:379 */
/* preorder tokens: leaf = protein index, internal node = -(number of children) then its children;
 * returns the token count (out may be NULL to size the buffer) */
uint64_t kc_tree_serialize(const kc_tree* t, int64_t* out, uint64_t capacity);
/* cluster_of[p] = index of the root child (top-level cluster) holding protein p */
int kc_tree_clusters(const kc_tree* t, uint32_t* cluster_of, uint32_t* n_clusters);

#ifdef __cplusplus
}
#endif
#endif /* KC_HOST_H_ */
