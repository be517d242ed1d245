//! Stand-alone wrapper crate around `b200.rs` so that the binding can be type-checked outside liblrge.
//! Inside liblrge the module is `crate::b200` and `Platform` / `LrgeError` are liblrge's own types
//! (liblrge/src/lib.rs:163-180, liblrge/src/error.rs:6-33); the small copies below stand in for them here.
pub mod b200;

#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub enum Platform {
    PacBio,
    Nanopore,
}

#[derive(Debug)]
pub enum LrgeError {
    FastqParseError(String),
    MapError(String),
    DuplicateReadIdentifier(String),
}

pub type Result<T> = std::result::Result<T, LrgeError>;
