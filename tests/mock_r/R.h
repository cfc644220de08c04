/* stand-in, see Rinternals.h */
