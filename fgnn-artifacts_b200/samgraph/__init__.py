"""samgraph — API-compatible front end of the B200-native sampling/extraction runtime.

Drop-in for the reference's `samgraph` package (SJTU-IPADS/fgnn-artifacts): the training scripts
under example/samgraph import `samgraph.torch as sam` and keep working unchanged."""
