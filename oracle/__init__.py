"""TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference algorithm for the
hot path (see oracle/jets_oracle.py).  Only tests/, __graft_entry__.smoke() and
bench.py's CPU-baseline / reference arm may import this package; the product
package ``emotivoice_b200`` never does."""
